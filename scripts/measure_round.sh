set -x
T=r1u
python bench.py --steps 50 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 5 --warmup 3 --only-iterations > gpurun_out/${T}_ncu_launch.log 2>&1
LRR_PROF_SETTLE=15 LRR_PROF_RANGE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_${T} -f python scripts/prof_iter.py 100000 1 > gpurun_out/${T}_ncu_full.log 2>&1
tail -3 gpurun_out/${T}_ncu_full.log
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json')); print(d['value'], d['value_l2_warm'], d['e2e']['value'], d['registrations']['value'], d['roofline']['kernel'], d['roofline']['frac'], {k:round(v['us_per_launch'],1) for k,v in d['kernels'].items()})"
