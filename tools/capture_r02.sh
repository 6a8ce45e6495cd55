#!/bin/bash
# Round-2 evidence in one gpurun call (1 GPU): parity tests, the bench line (with the CPU leg and the recursion-shaped
# proof), the CPU reference arm, ncu --set full of the NTT passes and of the leaf hash, the launch list of one bench step.
mkdir -p gpurun_out
TAG=${1:-r02}
export GL_REQUIRE_GPU=1
(time timeout 500 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_$TAG.log 2>&1
tail -3 gpurun_out/pytest_$TAG.log
timeout 400 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cut -c1-200 gpurun_out/bench_$TAG.json
timeout 200 python bench.py --no-cpu --no-extra --steps 3 --ntt-group 2147483648 > gpurun_out/bench_twocopy_$TAG.json 2> gpurun_out/bench_twocopy_$TAG.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
ncu --set full --clock-control none --import-source on -k regex:"k_ntt" -c 4 \
    -o gpurun_out/prof_ntt_$TAG -f python bench.py --steps 1 --warmup 1 --cols 64 --no-cpu --no-extra --no-ntt > gpurun_out/prof_ntt_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_leaf_hash" -s 1 -c 1 \
    -o gpurun_out/prof_leaf_$TAG -f python bench.py --steps 1 --warmup 1 --log-n 16 --no-cpu --no-ntt --no-extra > gpurun_out/prof_leaf_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/launches_$TAG.log 2>&1
ls -la gpurun_out | tail -8
