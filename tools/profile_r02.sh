#!/bin/bash
# Round-2 profile pass on one B200 (run through gpurun; writes gpurun_out/r02_*).  Every ncu command runs only
# after the same program has exited 0 without ncu (B200_PROFILING.md).
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1 || exit 1
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench.err
# launch list of the benchmark command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1
# both traceback flavours of config 2, one chunk each (same section set)
PAIRS=131072 STEPS=1 python tools/c2_flavours.py > gpurun_out/r02_flavours_small.jsonl 2>&1 || exit 1
python tools/c2_flavours.py > gpurun_out/r02_flavours.jsonl 2>&1
PAIRS=131072 STEPS=1 ncu --set full --import-source on --clock-control none -k regex:"psa_pack_fill_kernel|psa_pack_tb_kernel|psa_pack_rwalk_kernel" \
    -c 12 -o gpurun_out/r02_c2_kernels python tools/c2_flavours.py > gpurun_out/r02_ncu_c2.log 2>&1
# similarity kernel (f-4)
python tools/sim_bench.py > gpurun_out/r02_sim.log 2>&1
ncu --set full --clock-control none -k regex:psa_similarity -c 2 -o gpurun_out/r02_sim python tools/sim_bench.py > gpurun_out/r02_ncu_sim.log 2>&1
# systolic kernel (one GPU's share of the 8-GPU config-4 run: 977 strips) and the row-block kernel at 1 Mbp
M=200000 N=125000 python tools/prof_systolic.py > gpurun_out/r02_sys_plain.log 2>&1 || exit 1
M=200000 N=125000 ncu --set full --clock-control none -k regex:systolic_kernel -s 1 -c 1 -o gpurun_out/r02_systolic python tools/prof_systolic.py > gpurun_out/r02_ncu_sys.log 2>&1
ls -la gpurun_out | grep r02_
