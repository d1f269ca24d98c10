# A/B of k_objcov builds (development aid): cdcl_obj_fun_cov evaluation at config-5 size for every library under
# lrr_b200/csrc/build/variants (made with -DLRR_OBJ_THREADS / -DLRR_OBJ_MINB / -DLRR_OBJ_PIPE).  Usage on a B200: bash scripts/ab_objcov.sh
V=lrr_b200/csrc/build/variants
run() { echo "== $1"; shift; env "$@" timeout 120 python scripts/objcov_time.py 1000000 20 2>&1 | tail -1; }
run default X=1
for so in $V/*.so; do run $(basename $so .so) LRR_B200_LIB=$so; done
run default-again X=1
# the partial-row finalizer is shared with k_accumulate: fixed-point iteration of the 100k pair for the default build and the variants
iter() { echo "== iteration $1"; shift; env "$@" LRR_PROF_SETTLE=15 timeout 120 python scripts/prof_iter.py 100000 30 2>&1 | tail -3 | grep -v launches; }
iter default X=1
[ -e $V/old_head.so ] && iter old_head LRR_B200_LIB=$V/old_head.so
iter default-again X=1
