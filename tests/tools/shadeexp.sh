# A/B of k_wf_shade variants (built with -DWF_SHADE_PREFETCH=n -DWF_SHADE_MIN_BLOCKS=n into librtb200_<name>.so)
SC="cornell-box:64 materialball:64 materialball_glass:64 MaterialsScene:64 coffee:64 bathroom:32 soup22:4"
for v in "$@"; do echo "== variant $v"; RTB200_LIB=$PWD/raytracingrenderer_b200/librtb200_$v.so python tests/tools/perf_probe.py $SC; done
