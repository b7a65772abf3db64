timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for sc in materialball materialball_glass cornell-box coffee; do
  spp=256; [ $sc = coffee ] && spp=64; [ $sc = cornell-box ] && spp=64
  for r in 0 1; do
  echo "== reuse$r $sc"; RTB_REUSE=$r python tests/tools/profile_render.py $sc $spp 2>&1 | tail -2
  done
done
