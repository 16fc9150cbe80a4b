set -x
cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/bench_s3_n1.json 2> gpurun_out/bench_s3_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench_s3_n1.err
python tools/profile_factor.py && timeout 300 ncu --set full --clock-control none --import-source on -k regex:jacobi_cluster -c 1 -o gpurun_out/jacobi_cluster_r02 python tools/profile_factor.py > gpurun_out/ncu_jacobi_cluster.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_s3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_s3.log 2>&1
ls -la gpurun_out/
