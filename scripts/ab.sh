# A/B matrix (development aid): steady-state iteration time (L2 warm) and per-kernel CUDA-event times for environment switches
# and variant builds (scripts/build_variant.sh NAME -DMACRO=...).  Usage on a B200: bash scripts/ab.sh
V=lrr_b200/csrc/build/variants
run() { echo "== $1"; shift; env "$@" LRR_PROF_SETTLE=15 python scripts/prof_iter.py 100000 30 2>&1 | tail -3 | grep -v launches; }
run default X=1
run no-speculative-plan LRR_NO_SPEC=1
run no-norm-skip LRR_NO_SPEC=1 LRR_NO_NORM_SKIP=1
run no-mailbox LRR_NO_MAILBOX=1
run no-pdl LRR_NO_PDL=1
run no-certificates LRR_NO_PROOF=1
for so in $V/*.so; do [ -e "$so" ] && run "variant $(basename $so .so)" LRR_B200_LIB=$so; done
run default-again X=1
