#!/bin/bash
# Round-2 evidence pass, run ON the GPU box (one GPU):  gpurun -- bash tools/gather_r02.sh
# Everything lands in gpurun_out/r02/ ; tools/refresh_profiles.py turns it into the files committed under profiles/.
set -u
O=gpurun_out/r02; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt; nproc >> $O/gpu.txt; lscpu | grep -E "^CPU\(s\)|Model name" >> $O/gpu.txt
# 1. the default bench line (config 2, N = 1) with e2e and cpu baseline, and the reference arm
python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 > $O/bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > $O/bench_reference_arm.json
# 2. the other BASELINE configs
for c in 1 3 4 5; do python bench.py --config $c --steps 12 --warmup 3 2>/dev/null | tail -1; done > $O/configs.jsonl
# 3. ncu: launch list of the bench command, DRAM traffic of the full-workload launch, full captures of the hot kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:crb_tiled -s 1 -c 1 --csv --log-file $O/traffic.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:crb_tiled -s 3 -c 1 -f -o $O/ncu_stereo python bench.py --steps 2 --warmup 3 --streams 16 --seconds 240 --no-e2e --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:crb_tiled -s 3 -c 1 -f -o $O/ncu_mono python bench.py --config 4 --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:crb_tiled -s 1 -c 1 -f -o $O/ncu_8ch python tools/run_config.py 8 192000 44100 300 1 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:crb_tiled -s 1 -c 1 -f -o $O/ncu_sk python tools/run_config.py 2 48000 44100 60 16 3 > /dev/null 2>&1
# 4. where the warps' time goes (debug build with clock64 counters), if that build travelled
if [ -f variants/dbg/libclownresampler_b200.so ]; then
  for a in "64 600" "1024 10 1 22050 48000" "1 3600 8 192000 44100"; do CRB200_LIB=$PWD/variants/dbg/libclownresampler_b200.so python tools/dbg_timing.py $a; done > $O/warp_time.txt 2>&1
fi
# 5. microbenchmarks and the sanitizer pass
[ -x tools/microbench/pipes ] && tools/microbench/pipes > $O/pipe_microbench.jsonl 2>&1
[ -x tools/microbench/overlap ] && tools/microbench/overlap > $O/overlap.jsonl 2>&1
compute-sanitizer --tool memcheck python tools/sanitize_run.py > $O/sanitizer.txt 2>&1; tail -3 $O/sanitizer.txt
python tools/pcie_probe.py > $O/pcie_probe_n1.txt 2>&1
ls -la $O
