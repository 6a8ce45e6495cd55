#!/bin/bash
# Round-end evidence in ONE gpurun call (1 GPU, about 200 s on the box): parity tests, the bench line, the ncu launch
# list of a bench step and --set full captures of the dominant kernel and of the NTT passes, the microbenchmarks and
# the Poseidon variant ranking. Outputs in gpurun_out/; `python tools/update_profiles.py prof_leaf_final.ncu-rep` turns
# them into the tracked summaries under profiles/.
mkdir -p gpurun_out
(time timeout 300 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_final.log 2>&1
tail -4 gpurun_out/pytest_gpu_final.log
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
cut -c1-300 gpurun_out/bench_final.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-ntt --no-extra > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_leaf_hash" -s 1 -c 1 \
    -o gpurun_out/prof_leaf_final python bench.py --steps 1 --warmup 1 --log-n 16 --no-cpu --no-ntt --no-extra > gpurun_out/prof_leaf_final.log 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/microbench tools/microbench.cu > /dev/null 2>&1 && /tmp/microbench > gpurun_out/microbench.txt 2>&1
(cd tools/variants && ./out/variant_bench 234 out/*.cubin) > gpurun_out/variants_final.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_passA|k_passB" -s 40 -c 4 \
    -o gpurun_out/prof_ntt_final python bench.py --steps 1 --warmup 1 --cols 16 --no-cpu --no-extra > gpurun_out/prof_ntt_final.log 2>&1
ls -la gpurun_out | tail -12
