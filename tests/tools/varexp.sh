SC="coffee:64 bathroom:32 soup20:4 soup22:4"
for rep in 1 2; do for v in "" _t128; do
echo "== variant librtb200$v.so (run $rep)"; RTB200_LIB=$PWD/raytracingrenderer_b200/librtb200$v.so python tests/tools/perf_probe.py $SC
done; done
