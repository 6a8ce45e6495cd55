#!/bin/bash
# ncu captures for the hot kernels (run under gpurun; outputs in gpurun_out/). One GPU only.
set -x
mkdir -p gpurun_out
# 1) launch list of bench steps (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-ntt --no-extra > gpurun_out/launches_bench.log 2>&1
# 2) full capture of the NTT passes at the 2^20 shape (16 columns)
ncu --set full --clock-control none --import-source on -k regex:"k_passA|k_passB" -s 40 -c 4 \
    -o gpurun_out/prof_ntt python bench.py --steps 1 --warmup 1 --cols 16 --no-cpu --no-extra > gpurun_out/prof_ntt.log 2>&1
# 3) full capture of the Poseidon leaf hash (2^19 leaves x 234) and one Merkle level
ncu --set full --clock-control none --import-source on -k regex:"k_leaf_hash" -s 1 -c 1 \
    -o gpurun_out/prof_leaf python bench.py --steps 1 --warmup 1 --log-n 16 --no-cpu --no-ntt --no-extra > gpurun_out/prof_leaf.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_merkle_level" -s 1 -c 1 \
    -o gpurun_out/prof_level python bench.py --steps 1 --warmup 1 --log-n 16 --no-cpu --no-ntt --no-extra > gpurun_out/prof_level.log 2>&1
ls -la gpurun_out
