V=lrr_b200/csrc/build/variants
run() { echo "== $1"; shift; env "$@" LRR_PROF_SETTLE=15 python scripts/prof_iter.py 100000 30 2>&1 | tail -3 | grep -v launches; }
run head LRR_B200_LIB=$V/head.so
run current X=1
run noalt LRR_NO_ALT=1
run accflat LRR_ACC_FLAT=1
run solvenoinline LRR_B200_LIB=$V/solvenoinline.so
run solvenoinline+flat LRR_B200_LIB=$V/solvenoinline.so LRR_ACC_FLAT=1
run mq LRR_MQ=1
run nblk2 LRR_NBLK_PER_SM=2
run nblk1 LRR_NBLK_PER_SM=1
run current-again X=1
