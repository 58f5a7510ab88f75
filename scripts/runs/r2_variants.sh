timeout 600 python -m pytest tests/test_refine.py -x -q 2>&1 | tail -30
