# one measurement round on the GPU box: bash scripts/measure_round.sh TAG   (outputs under gpurun_out/TAG_*; gpurun copies back at most
# 64 MiB: only the fixed-point capture imports the sources)
set -x
T=${1:-r2a}
python bench.py --steps 50 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 5 --warmup 3 --only-iterations > gpurun_out/${T}_ncu_launch.log 2>&1
# in-registration finest-level iterations (2nd onwards: certificates in place, transform still moving) — the state `value` is quoted on
LRR_PROF_INREG=2 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/prof_${T}_inreg -f python scripts/prof_iter.py 100000 1 > gpurun_out/${T}_ncu_inreg.log 2>&1
# the FIRST finest-level iteration of the registration (no certificates yet: the whole find_closest the bench line's `roofline` names)
LRR_PROF_INREG=1 ncu --set full --clock-control none --profile-from-start off -c 10 -o gpurun_out/prof_${T}_first -f python scripts/prof_iter.py 100000 1 > gpurun_out/${T}_ncu_first.log 2>&1
# one fixed-point iteration (round 1's state, kept for comparison)
LRR_PROF_SETTLE=15 LRR_PROF_RANGE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_${T} -f python scripts/prof_iter.py 100000 1 > gpurun_out/${T}_ncu_full.log 2>&1
tail -3 gpurun_out/${T}_ncu_full.log gpurun_out/${T}_ncu_inreg.log
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json')); print(d['value'], d['value_l2_warm'], d['e2e']['value'], d['registrations']['value'], d['roofline']['kernel'], d['roofline']['frac'], {k:round(v['us_per_launch'],1) for k,v in d['kernels'].items()})"
