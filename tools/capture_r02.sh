#!/bin/bash
# Round-2 evidence in one gpurun call: ncu --set full of the NTT passes (iNTT col/row-natural, LDE col/row-bitrev)
# and of the leaf hash, plus the launch list of one bench step.
mkdir -p gpurun_out
TAG=${1:-r02}
ncu --set full --clock-control none --import-source on -k regex:"k_ntt" -c 4 \
    -o gpurun_out/prof_ntt_$TAG -f python bench.py --steps 1 --warmup 1 --cols 64 --no-cpu --no-extra --no-ntt > gpurun_out/prof_ntt_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_leaf_hash" -s 1 -c 1 \
    -o gpurun_out/prof_leaf_$TAG -f python bench.py --steps 1 --warmup 1 --log-n 16 --no-cpu --no-ntt --no-extra > gpurun_out/prof_leaf_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/launches_$TAG.log 2>&1
ls -la gpurun_out | tail -8
