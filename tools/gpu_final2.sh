#!/bin/bash
# Round-end evidence run, part 2: ncu launch list of the bench command + one full capture of the tensor-core kernels.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_final.log 2>&1
echo "ncu launches exit $?" > gpurun_out/summary_final2.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_halo|conv1_umma" -s 48 -c 12 -o gpurun_out/prof_final -f python bench.py --steps 6 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/ncu_full_final.log 2>&1
echo "ncu full exit $?" >> gpurun_out/summary_final2.txt
cat gpurun_out/summary_final2.txt; ls -la gpurun_out/prof_final.ncu-rep; tail -3 gpurun_out/ncu_full_final.log
