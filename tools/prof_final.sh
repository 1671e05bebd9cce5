# One GPU call that regenerates everything under profiles/ for the current build (see profiles/README.md).
set -x
mkdir -p gpurun_out
DISPNET_B200_GRAPHS=0 DISPNET_B200_SIDE_STREAM=0 DN_PDL=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --minimal > gpurun_out/launches_final.log 2>&1
tail -2 gpurun_out/launches_final.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:igemm_tc -s 2 -c 1 -f -o gpurun_out/full_igemm_tc_feat27 python tools/prof_conv.py feat27 > gpurun_out/full1.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:igemm_halo -s 2 -c 1 -f -o gpurun_out/full_igemm_halo_feat3 python tools/prof_conv.py feat3 > gpurun_out/full2.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:wgrad_tc -s 2 -c 1 -f -o gpurun_out/full_wgrad_tc_feat27 python tools/prof_conv.py feat27 > gpurun_out/full3.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:bnf_bwd_reduce -s 12 -c 1 -f -o gpurun_out/full_bnf_bwd_reduce python bench.py --steps 1 --warmup 3 --minimal > gpurun_out/full4.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:bnf_apply -s 0 -c 1 -f -o gpurun_out/full_bnf_apply python bench.py --steps 1 --warmup 3 --minimal > gpurun_out/full5.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:fwd_bulk -s 1 -c 1 -f -o gpurun_out/full_hc_fwd python tools/prof_head.py > gpurun_out/full6.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 600 python bench.py --dump gpurun_out/layers_final.txt > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log
timeout 200 python tools/prof_bn.py > gpurun_out/prof_bn.txt 2>&1
timeout 100 python tools/prof_head.py > gpurun_out/prof_head.txt 2>&1
