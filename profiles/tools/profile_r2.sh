#!/bin/bash
# round-2 captures: launch list of the bench command, then --set full of one launch of every hot kernel
T=${1:-r2z}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_launches.log 2>&1
K2AB_REPS=2 python profiles/tools/k2ab.py prof > gpurun_out/${T}_k2ab_plain.log 2>&1 &&
K2AB_REPS=2 ncu --set full --clock-control none --import-source on -k regex:"k_emit|k_plan" -c 16 -o gpurun_out/${T}_emit python profiles/tools/k2ab.py prof > gpurun_out/${T}_ncu_emit.log 2>&1
REPS=2 LAGS=0 python profiles/tools/multi.py > gpurun_out/${T}_multi_plain.log 2>&1 &&
REPS=2 LAGS=0 ncu --set full --clock-control none --import-source on -k regex:"k_emit_multi|k_multi_order" -s 0 -c 3 -o gpurun_out/${T}_multi python profiles/tools/multi.py > gpurun_out/${T}_ncu_multi.log 2>&1
SIX_REPS=2 python profiles/tools/six.py > gpurun_out/${T}_six_plain.log 2>&1 &&
SIX_REPS=2 ncu --set full --clock-control none --import-source on -k regex:"k_six|k_stop" -c 12 -o gpurun_out/${T}_six python profiles/tools/six.py > gpurun_out/${T}_ncu_six.log 2>&1
ls -la gpurun_out/${T}_*
