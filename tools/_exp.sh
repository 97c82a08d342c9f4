for d in 0 1 2 8 10 4 14 15; do echo "== E2E_TC_DEBUG=$d"; E2E_TC_DEBUG=$d timeout 120 python tools/bench_layers.py --only "loc" --ops wgrad 2>&1 | grep "loc4\|loc3\|loc2"; done
