#!/bin/bash
# ncu captures for the hot kernels (run under gpurun; outputs in gpurun_out/). One GPU only.
set -x
mkdir -p gpurun_out
# 1) launch list of one bench step region (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-ntt > gpurun_out/launches_bench.log 2>&1
# 2) full capture of the NTT passes at the 2^20 shape (16 columns, 2 groups) and the bare NTT
ncu --set full --clock-control none --import-source on -k regex:"k_passA|k_passB" -s 40 -c 6 \
    -o gpurun_out/prof_ntt python bench.py --steps 1 --warmup 1 --cols 16 --no-cpu > gpurun_out/prof_ntt.log 2>&1
# 3) full capture of the Poseidon leaf hash at 2^19 leaves x 234 and a Merkle level
ncu --set full --clock-control none --import-source on -k regex:"k_leaf_hash|k_merkle_level" -s 2 -c 3 \
    -o gpurun_out/prof_hash python bench.py --steps 1 --warmup 1 --log-n 16 --no-cpu --no-ntt > gpurun_out/prof_hash.log 2>&1
ls -la gpurun_out
