mkdir -p gpurun_out/r4i
for v in 2 1; do
MFG_STEP_KERNEL=$v timeout 60 python bench.py --steps 300 --warmup 20 --age 300 --no-cpu --no-e2e > gpurun_out/r4i/b_$v.json 2> gpurun_out/r4i/b_$v.err
done
python tools/bench_brief.py gpurun_out/r4i/b_*.json
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "freerun or reset or replay" > gpurun_out/r4i/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4i/pytest.log
tail -2 gpurun_out/r4i/pytest.log
