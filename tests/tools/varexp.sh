SC="cornell-box:64 materialball:64 MaterialsScene:64 coffee:64 bathroom:32 soup20:4 soup22:4"
for ps in 0 1; do
echo "== RTB_SHADOW_PERSISTENT=$ps"; RTB_SHADOW_PERSISTENT=$ps python tests/tools/perf_probe.py $SC
done
