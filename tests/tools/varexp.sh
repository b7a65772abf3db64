SC="cornell-box:64 materialball:64 coffee:64 bathroom:32 soup20:4 soup22:4 soup24:4"
echo "== host SAH builder (RTB_GPU_BUILD=0)"; RTB_GPU_BUILD=0 python tests/tools/perf_probe.py $SC
echo "== device LBVH builder (RTB_GPU_BUILD=1)"; RTB_GPU_BUILD=1 python tests/tools/perf_probe.py $SC
