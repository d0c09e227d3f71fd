for i in 1 2 3; do CHAIN_TIMING_MIN=1000000 timeout 40 python scripts/chain_check.py 10000 2>&1 | grep -E "^n=" | tail -1; done
for i in 1 2; do timeout 60 python scripts/chain_check.py 20000 2>&1 | grep -E "^n=|chain=True" | tail -2; echo rc=$?; done
for i in 1 2; do timeout 60 python scripts/chain_check.py 50000 2>&1 | grep -E "^n=|chain=True|stage" | tail -3; echo rc=$?; done
