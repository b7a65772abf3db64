set -x
export RTB_REUSE=0
ncu --set full --clock-control none --import-source on -k regex:k_wf -s 60 -c 3 -o gpurun_out/r02_final_materialball -f python tests/tools/profile_render.py materialball 64
RTB_SHADOW_PERSISTENT=1 ncu --set full --clock-control none -k regex:k_wf -s 62 -c 3 -o /tmp/r02_final_bathroom -f python tests/tools/profile_render.py bathroom 32
RTB_SHADOW_PERSISTENT=1 ncu --set full --clock-control none -k regex:k_wf -s 8 -c 3 -o /tmp/r02_final_soup22 -f python tests/tools/profile_soup.py 22 4
python tests/tools/ncu_summary.py gpurun_out/r02_final_materialball.ncu-rep > gpurun_out/r02_final_materialball_summary.md
python tests/tools/ncu_summary.py /tmp/r02_final_bathroom.ncu-rep > gpurun_out/r02_final_bathroom_summary.md
python tests/tools/ncu_summary.py /tmp/r02_final_soup22.ncu-rep > gpurun_out/r02_final_soup22_summary.md
python bench.py --steps 1 --warmup 3 --spp 16 --e2e-steps 1 --no-cpu --no-per-scene --no-strong > gpurun_out/r02_bench_spp16.json && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_bench_spp16_launches.csv python bench.py --steps 1 --warmup 3 --spp 16 --e2e-steps 1 --no-cpu --no-per-scene --no-strong > gpurun_out/r02_bench_spp16_under_ncu.json
python tests/tools/ncu_launches.py gpurun_out/r02_bench_spp16_launches.csv > gpurun_out/r02_bench_spp16_summary.txt
du -sh gpurun_out
