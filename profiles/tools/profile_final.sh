#!/bin/bash
# final round-2 captures (short: GPU budget): bench line, launch list of our kernels, --set full of the changed kernels
T=${1:-r2y}
timeout 200 python bench.py > gpurun_out/${T}_bench1.log 2>&1
timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_plain.log 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_launches.log 2>&1
K2AB_REPS=1 timeout 100 python profiles/tools/k2ab.py prof > gpurun_out/${T}_k2ab_plain.log 2>&1 &&
K2AB_REPS=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"^k_emit_nuc$|^k_emit_prot$|^k_plan_pieces$|^k_plan_records$" -c 8 -o gpurun_out/${T}_emit python profiles/tools/k2ab.py prof > gpurun_out/${T}_ncu_emit.log 2>&1
SIX_REPS=2 timeout 100 python profiles/tools/six.py > gpurun_out/${T}_six_plain.log 2>&1 &&
SIX_REPS=2 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_six_aa|k_six_cand|k_six_bits_ix" -s 3 -c 3 -o gpurun_out/${T}_six python profiles/tools/six.py > gpurun_out/${T}_ncu_six.log 2>&1
ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
tail -1 gpurun_out/${T}_bench1.log | cut -c1-300
