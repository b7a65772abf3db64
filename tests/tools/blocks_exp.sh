SC="cornell-box:64 materialball:128 MaterialsScene:128 coffee:64 bathroom:32"
for b in 8 7 6 5 4; do echo "== RTB_EXTEND_BLOCKS_PER_SM=$b"; RTB_EXTEND_BLOCKS_PER_SM=$b python tests/tools/perf_probe.py $SC; done
