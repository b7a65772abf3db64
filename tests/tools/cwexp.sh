SC="cornell-box:64 materialball:64 MaterialsScene:64 coffee:64 bathroom:32 soup20:4"
echo "== FAST"; python tests/tools/perf_probe.py $SC
for kb in 0 32; do for sh in 1 0; do
echo "== CW stage=${kb}KB persistent_shadow=$sh"; RTB_CW_STAGE_KB=$kb RTB_CW_SHADOW=$sh python tests/tools/perf_probe.py --trav cw $SC
done; done
