# A/B of library builds: tests/tools/abexp.sh "<scene specs>" <lib suffix> <lib suffix> ...   ("" = the shipped librtb200.so)
SC="$1"; shift
for v in "$@"; do
  for rep in 1 2; do
    if [ -z "$v" ] || [ "$v" = "shipped" ]; then echo "== shipped ($rep)"; python tests/tools/perf_probe.py $SC
    else echo "== variant $v ($rep)"; RTB200_LIB=$PWD/raytracingrenderer_b200/librtb200_$v.so python tests/tools/perf_probe.py $SC; fi
  done
done
