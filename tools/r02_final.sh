#!/bin/bash
# Final single-GPU round trip of round 2: GPU tests, smoke, bench (+ reference arm), sweep, launch list, ncu captures of the
# training (2^22, both input forms), inference and encode kernels.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.txt
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --steps 100 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
python tools/sweep.py > gpurun_out/sweep_1gpu.txt 2>&1; tail -22 gpurun_out/sweep_1gpu.txt
python tools/lab_train.py one 2>&1 | tail -1 | tee gpurun_out/lab_train_final.txt
python tools/bench_records.py 2>&1 | tail -2 | tee -a gpurun_out/lab_train_final.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nrc_train_kernel -s 2 -c 1 -o gpurun_out/r02f_train_2p22_unpacked -f python tools/prof_train.py 22 unpacked 2 > gpurun_out/ncu_t1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nrc_train_kernel -s 2 -c 1 -o gpurun_out/r02f_train_2p22_encoded -f python tools/prof_train.py 22 encoded 2 > gpurun_out/ncu_t2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nrc_infer_kernel -s 3 -c 1 -o gpurun_out/r02f_infer_enc -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_i1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nrc_encode_kernel -s 2 -c 1 -o gpurun_out/r02f_encode -f python tools/prof_encode.py > gpurun_out/ncu_e1.log 2>&1
python bench.py --impl reference --steps 100 --warmup 5 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; head -c 400 gpurun_out/bench_reference.json
ls -la gpurun_out/*.ncu-rep
