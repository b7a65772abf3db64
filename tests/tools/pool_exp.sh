SC="materialball:256 materialball_glass:256 coffee:64"
for s in 8388608 12582912 16777216; do for k in 2 3 4; do echo "== RTB_POOL_SLOTS=$s RTB_POOLS=$k"; RTB_POOL_SLOTS=$s RTB_POOLS=$k python tests/tools/perf_probe.py $SC; done; done
