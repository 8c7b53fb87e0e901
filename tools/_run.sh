python -m pytest tests/test_gpu_parity.py tests/test_cli_gpu.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload config3 --steps 100 --warmup 5 --no-e2e --no-cpu --cc-sort
python tools/sort_probe.py 2>&1 | tail -3
