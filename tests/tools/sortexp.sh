set -x
SC="cornell-box:64 materialball:64 MaterialsScene:64 coffee:64 bathroom:32 soup20:4"
for cfg in "0 0" "1 0" "0 1" "1 1"; do
  set -- $cfg
  echo "== SORT_SHADOW=$1 SORT_EXTEND=$2"
  RTB_SORT_SHADOW=$1 RTB_SORT_EXTEND=$2 python tests/tools/perf_probe.py $SC
done
