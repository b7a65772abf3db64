# RTB_TREE_OPT=<passes>[:fraction of the nodes, by area] x child order for any-hit rays
SC="cornell-box:64 materialball:64 materialball_glass:64 MaterialsScene:64 coffee:64 bathroom:32 soup20:4"
run() { echo "== $1"; shift; env "$@" python tests/tools/perf_probe.py $SC 2>&1; }
run "baseline" A=0
run "order 3 (p/C, SAH cost)" RTB_TREE_ORDER=3
run "order 4 (p/C, log cost)" RTB_TREE_ORDER=4
run "opt 2 x 5% + order 3" RTB_TREE_OPT=2:0.05 RTB_TREE_ORDER=3
run "opt 2 x 5% + order 4" RTB_TREE_OPT=2:0.05 RTB_TREE_ORDER=4
