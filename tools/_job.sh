mkdir -p gpurun_out/r4m
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r4m/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4m/pytest.log
tail -2 gpurun_out/r4m/pytest.log
timeout 60 python bench.py --steps 300 --warmup 20 --age 300 --no-cpu --no-e2e > gpurun_out/r4m/b.json 2> gpurun_out/r4m/b.err
python tools/bench_brief.py gpurun_out/r4m/b.json
