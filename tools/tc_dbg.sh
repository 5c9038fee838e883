for d in 0 1 2 4 8 3 12 15; do echo "== SPG_TC_DBG=$d"; SPG_TC_DBG=$d timeout 100 python tools/tc_bench.py 482304 2>&1 | sed -n 4,5p | cut -c1-110; done
