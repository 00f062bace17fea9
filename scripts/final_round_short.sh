t=r01d
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${t}_tests_gpu.log 2>&1; tail -1 gpurun_out/${t}_tests_gpu.log
timeout 400 python bench.py > gpurun_out/${t}_bench_cfg3.json 2> gpurun_out/${t}_bench_cfg3.err
for w in cfg2 cfg5; do timeout 300 python bench.py --workload $w > gpurun_out/${t}_bench_$w.json 2>/dev/null; done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${t}_smoke.log 2>&1; tail -1 gpurun_out/${t}_smoke.log
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-train > gpurun_out/${t}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${t}_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-train > gpurun_out/${t}_ncu_launch.log 2>&1
