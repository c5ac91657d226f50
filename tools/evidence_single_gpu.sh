# Single-GPU evidence batch of a round (run under gpurun; outputs land in gpurun_out/, the .ncu-rep files are summarised
# to CSV on the box and removed: gpurun merges at most 64 MiB back)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_r2_config4_n1.json 2> gpurun_out/bench_c4.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_config4_reference_arm.json 2> gpurun_out/bench_ref.err
for c in 1 2 3 5; do python bench.py --config $c > gpurun_out/bench_r2_config${c}_n1.json 2> gpurun_out/bench_c$c.err; done
python tools/pass_layers.py > gpurun_out/conv_layers_r2_fastdvdnet_pass.txt 2>&1
python tools/ffdnet_pass.py > gpurun_out/ffdnet_pass_r2.txt 2>&1
SCI_FFDNET_INF=tf32 python tools/ffdnet_pass.py >> gpurun_out/ffdnet_pass_r2.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_stage2_fastdvd_512x512x8.csv python tools/run_stage2.py 512 21,2 > gpurun_out/run_stage2.log 2>&1
ncu --set full --clock-control none -k regex:conv_fwd -s 34 -c 17 -f -o /tmp/prof_half python tools/run_fwd.py 2 > gpurun_out/ncu_half.log 2>&1
python tools/ncu_summary.py /tmp/prof_half.ncu-rep gpurun_out/ncu_r2_conv_fwd_half_pass.csv > /dev/null
ncu --set full --clock-control none -k regex:conv_fwd2 -c 12 -f -o /tmp/prof_ffd python tools/ffdnet_pass.py > gpurun_out/ncu_ffd.log 2>&1
python tools/ncu_summary.py /tmp/prof_ffd.ncu-rep gpurun_out/ncu_r2_conv_fwd2_ffdnet_split.csv > /dev/null
du -sh gpurun_out
