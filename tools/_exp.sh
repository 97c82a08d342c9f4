for d in 0 1 2 0 1; do echo "== E2E_TC_POLL=$d"; E2E_TC_POLL=$d timeout 120 python tools/bench_layers.py --only "loc" --ops fwd,dgrad 2>&1 | grep "loc4\|loc3\|loc2\|weighted"; done
