python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -12
python -m pytest tests -m gpu -x -q -k "unlagged" 2>&1 | tail -3
