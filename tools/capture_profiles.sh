#!/bin/bash
# Round evidence in one GPU call (run through gpurun from the repo root): bench records of the three configurations and the
# reference arm, the ncu launch list of a training step, DRAM traffic of the conv / wgrad launches, and --set full captures of
# the first convolution launches of a step.  Outputs go to gpurun_out/<tag>_*; tools/summarize_ncu.py and
# tools/measure_traffic.py turn them into the text files under profiles/.
set -u
TAG=${1:-r02z}
O=gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_conv_tma|k_wgrad_tma' --csv --log-file $O/${TAG}_traffic.csv python tools/prof_step.py 2 > $O/${TAG}_traffic.log 2>&1
python tools/measure_traffic.py $O/${TAG}_traffic.csv profiles/traffic.json "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch over two training steps (tools/prof_step.py 2), captured in the same gpurun call as the bench record ${TAG}" > $O/${TAG}_traffic.txt
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
python bench.py --config 2 --steps 20 --warmup 5 > $O/${TAG}_bench_config2.json 2>> $O/${TAG}_bench.err
python bench.py --config 5 --steps 10 --warmup 3 > $O/${TAG}_bench_config5.json 2>> $O/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches.csv python tools/prof_step.py 2 > $O/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_conv_tma -s 64 -c 4 -o $O/${TAG}_conv -f python tools/prof_step.py 2 > $O/${TAG}_conv.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_wgrad_tma -s 32 -c 2 -o $O/${TAG}_wgrad -f python tools/prof_step.py 2 > $O/${TAG}_wgrad.log 2>&1
ls -la $O/${TAG}_*
