python scripts/batch_cost_dump.py 2>&1 | tail -3
python scripts/cdc_time.py 1000000 300 2>&1 | tail -4
python -m pytest tests/test_cdc.py -m gpu -x -q 2>&1 | tail -3
