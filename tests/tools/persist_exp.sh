SC="coffee:64 bathroom:32 soup20:4"
for v in auto 0 1; do echo "== RTB_SHADOW_PERSISTENT=$v"; if [ $v = auto ]; then python tests/tools/perf_probe.py $SC; else RTB_SHADOW_PERSISTENT=$v python tests/tools/perf_probe.py $SC; fi; done
