timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sparse or C5 or EXP" 2>&1 | tail -12
