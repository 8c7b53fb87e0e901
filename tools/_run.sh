python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -m gpu -x -q 2>&1 | tail -3
for wl in config2 config4 config3; do
python bench.py --workload $wl --steps 50 --warmup 3 --no-e2e --no-cpu 
done
