SC="cornell-box:64 materialball:64 MaterialsScene:64 coffee:64 bathroom:32 soup20:4 soup22:4"
echo "== FAST, shared stack 12"; python tests/tools/perf_probe.py $SC
echo "== FAST, local stack"; RTB200_LIB=$PWD/raytracingrenderer_b200/librtb200_ss0.so python tests/tools/perf_probe.py $SC
echo "== Q16, shared stack 12"; python tests/tools/perf_probe.py --trav q16 $SC
echo "== Q16, local stack"; RTB200_LIB=$PWD/raytracingrenderer_b200/librtb200_ss0.so python tests/tools/perf_probe.py --trav q16 $SC
