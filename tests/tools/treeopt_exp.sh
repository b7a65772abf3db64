# RTB_TREE_OPT=<passes>[:fraction of the nodes, by area] x child order for any-hit rays
SC="cornell-box:64 materialball:64 MaterialsScene:64 coffee:64 bathroom:32 soup20:4"
run() { echo "== $1"; shift; env "$@" python tests/tools/perf_probe.py $SC 2>&1; }
run "baseline" A=0
run "smaller child first" RTB_TREE_ORDER=2
run "opt 2 x 5% + smaller child first" RTB_TREE_OPT=2:0.05 RTB_TREE_ORDER=2
run "opt 4 x 5% + smaller child first" RTB_TREE_OPT=4:0.05 RTB_TREE_ORDER=2
