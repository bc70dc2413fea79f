timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; cut -c1-330 gpurun_out/bench_final.json
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python tools/bench_kernels.py > gpurun_out/kernels_final.json 2> gpurun_out/kernels_final.err; tail -2 gpurun_out/kernels_final.err
