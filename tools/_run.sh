python -m pytest tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -2
for wl in config2 config4; do
echo "== $wl one call"; python bench.py --workload $wl --steps 50 --warmup 3 --no-e2e --no-cpu --fused | cut -c1-120
echo "== $wl three calls"; python bench.py --workload $wl --steps 50 --warmup 3 --no-e2e --no-cpu | cut -c1-120
done
