# round-2 evidence session for the classifier kernels of the last commits (run under gpurun from the repo root)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ncu_r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-pageable --no-config2 --no-extras > gpurun_out/ncu_bench.log 2>&1; echo ncu-list rc=$?
python tools/prof_s2v.py bipartite > gpurun_out/s2v_timing_r02.txt 2>&1
python tools/prof_s2v.py complete >> gpurun_out/s2v_timing_r02.txt 2>&1
python tools/prof_s2v_grad.py >> gpurun_out/s2v_timing_r02.txt 2>&1
python tools/time_gram.py >> gpurun_out/s2v_timing_r02.txt 2>&1
python tools/prof_s2v_general.py >> gpurun_out/s2v_timing_r02.txt 2>&1
cat gpurun_out/s2v_timing_r02.txt
timeout 300 $NCU -k regex:s2v_gram_tc_kernel -s 2 -c 1 -o gpurun_out/ncu_r02_gram python tools/prof_s2v.py complete > gpurun_out/ncu_gram.log 2>&1; echo gram rc=$?
timeout 300 $NCU -k regex:s2v_complete_kernel -s 2 -c 1 -o gpurun_out/ncu_r02_complete python tools/prof_s2v.py complete > gpurun_out/ncu_complete.log 2>&1; echo complete rc=$?
timeout 300 $NCU -k regex:s2v_bipartite_dense_kernel -s 2 -c 1 -o gpurun_out/ncu_r02_dense python tools/prof_s2v.py bipartite > gpurun_out/ncu_dense.log 2>&1; echo dense rc=$?
timeout 300 $NCU -k regex:s2v_bipartite_grad_kernel -s 2 -c 1 -o gpurun_out/ncu_r02_grad python tools/prof_s2v_grad.py > gpurun_out/ncu_grad.log 2>&1; echo grad rc=$?
timeout 300 $NCU -k regex:s2v_complete_grad_kernel -s 2 -c 1 -o gpurun_out/ncu_r02_cgrad python tools/time_gram.py > gpurun_out/ncu_cgrad.log 2>&1; echo cgrad rc=$?
timeout 300 $NCU -k regex:s2v_bipartite_general_grad_kernel -s 2 -c 1 -o gpurun_out/ncu_r02_ggrad python tools/prof_s2v_general.py > gpurun_out/ncu_ggrad.log 2>&1; echo ggrad rc=$?
ls -la gpurun_out/*.ncu-rep
