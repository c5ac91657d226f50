# Closing check of a round on one GPU: tests, smoke, the default bench line, config 5, and the ncu launch list of bench.py
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_r2_config4_n1.json 2> gpurun_out/bench_c4.err
python bench.py --config 5 > gpurun_out/bench_r2_config5_n1.json 2> gpurun_out/bench_c5.err
python bench.py --config 3 --no-gpu-eager --no-cpu-baseline > gpurun_out/bench_c3_quick.json 2> gpurun_out/bench_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_bench.csv python bench.py --steps 2 --warmup 1 --no-gpu-eager --no-cpu-baseline --no-delta-psnr > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_ncu.err
du -sh gpurun_out
