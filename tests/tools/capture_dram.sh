export RTB_REUSE=0
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum"
ncu --metrics $M --clock-control none -k regex:k_wf -s 60 -c 3 --csv --log-file gpurun_out/r02_dram_materialball.csv python tests/tools/profile_render.py materialball 64 > /dev/null
RTB_SHADOW_PERSISTENT=1 ncu --metrics $M --clock-control none -k regex:k_wf -s 62 -c 3 --csv --log-file gpurun_out/r02_dram_bathroom.csv python tests/tools/profile_render.py bathroom 32 > /dev/null
RTB_SHADOW_PERSISTENT=1 ncu --metrics $M --clock-control none -k regex:k_wf -s 8 -c 3 --csv --log-file gpurun_out/r02_dram_soup22.csv python tests/tools/profile_soup.py 22 4 > /dev/null
RTB_SHADOW_PERSISTENT=1 ncu --metrics $M --clock-control none -k regex:k_wf -s 8 -c 3 --csv --log-file gpurun_out/r02_dram_soup24.csv python tests/tools/profile_soup.py 24 4 > /dev/null
