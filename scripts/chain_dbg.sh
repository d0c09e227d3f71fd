CHAIN_TIMING_MIN=1000000 timeout 40 python scripts/chain_check.py 10000 2>&1 | grep -E "^n=" | tail -1
timeout 60 python scripts/chain_check.py 50000 2>&1 | grep -E "^n=|chain=True|stage" | tail -3
