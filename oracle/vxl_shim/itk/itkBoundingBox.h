// included by the reference but unused on the hot path (stand-in, test infrastructure only)
