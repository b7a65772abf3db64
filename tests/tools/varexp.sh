SC="materialball:64 coffee:64 bathroom:32 soup20:4"
for v in "" _ri8 _ri16 _ch64 _ch256; do
echo "== variant librtb200$v.so"; RTB200_LIB=$PWD/raytracingrenderer_b200/librtb200$v.so python tests/tools/perf_probe.py $SC
done
