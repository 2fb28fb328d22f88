mkdir -p gpurun_out
DDB_ROWREG_HYBRID=1 timeout 120 tools/row_timing 200 100 32768 3 > gpurun_out/rt_final.log 2>&1; tail -4 gpurun_out/rt_final.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu_r02.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; echo bench rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-pageable --no-config2 --no-extras > gpurun_out/ncu_bench.log 2>&1; echo ncu-list rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:simplex_rowreg_kernel -c 1 -f -o gpurun_out/ncu_r02_rowreg tools/row_timing 200 100 4662 1 > gpurun_out/ncu_r02_rowreg.log 2>&1; echo ncu-rowreg rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:simplex_cluster_kernel -c 1 -f -o gpurun_out/ncu_r02_cluster tools/cluster_timing 500 250 294 1 > gpurun_out/ncu_r02_cluster.log 2>&1; echo ncu-cluster rc=$?
timeout 300 python tools/time_shapes.py > gpurun_out/solve_shapes_r02.txt 2>&1; cat gpurun_out/solve_shapes_r02.txt
timeout 300 python tools/time_cluster.py > gpurun_out/cluster_vs_plan2_r02.jsonl 2>&1
timeout 600 python tools/run_configs.py > gpurun_out/configs_r02.jsonl 2>&1; echo configs rc=$?
python tools/fused_ab.py > gpurun_out/fused_ab_r02.jsonl; DDB_FUSED_INKERNEL=1 python tools/fused_ab.py >> gpurun_out/fused_ab_r02.jsonl; cat gpurun_out/fused_ab_r02.jsonl
