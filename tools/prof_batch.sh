set -x
timeout 600 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
timeout 200 python tools/ncu_step.py > gpurun_out/ncu_step_plain_f.log 2>&1 && timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_kernels_r1_final.csv python tools/ncu_step.py > gpurun_out/ncu_step_f.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short_plain.json 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_bench_r1_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_f.log 2>&1
timeout 200 python tools/ncu_bn.py > /dev/null 2>&1 && timeout 600 ncu --set full --import-source on --clock-control none -k regex:bn_.*fused -o gpurun_out/bn_fused_r1_final python tools/ncu_bn.py > gpurun_out/ncu_bn_f.log 2>&1
timeout 200 python tools/microbench.py --only fps_4096 > /dev/null 2>&1 && timeout 600 ncu --set full --import-source on --clock-control none -k regex:fps_pair -c 2 -o gpurun_out/fps_pair_r1_final python tools/microbench.py --only fps_4096 --iters 2 > gpurun_out/ncu_fps_f.log 2>&1
timeout 600 python tools/bench_configs.py > gpurun_out/configs_r1_final.log 2>&1
tail -3 gpurun_out/configs_r1_final.log
