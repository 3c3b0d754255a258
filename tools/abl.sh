for c in 0 2; do echo "cfg=$c"; CLO_RADIX_CFG=$c timeout 120 python tools/quick_bench.py --what sort,sortprof 2>&1 | grep -o '"ms": [0-9.]*\|"cycles_per_tile": {[^}]*}\|"walk_rounds": [0-9.]*\|"prop_.*' ; done
CLO_RADIX_CFG=2 timeout 120 python tools/check_v6.py | grep "^32" | tail -4
CLO_RADIX_CFG=2 CLO_RADIX_PP_FLAGS=8 timeout 120 python tools/check_v6.py | grep "^32" | tail -2
