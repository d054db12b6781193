set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
N=$(python bench.py --profile-launch smooth0 2>&1 | grep -o "[0-9]* igemm launches" | cut -d' ' -f1)
echo "igemm launches per forward: $N"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph"
$B > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"igemm_kernel|asf_|db_head|maxpool|stem_pack|pp_" -s 460 -c 200 --csv --log-file gpurun_out/launches_r1_share.csv $B > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
for L in smooth0 layer3.1.dcn layer2.1.dcn layer1.1.conv2 head.0 layer1.1.conv3 stem layer2.1.offset; do
  B="python bench.py --profile-launch $L"
  $B > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s $N -c 1 -f -o gpurun_out/prof4_$L $B > gpurun_out/ncu4_$L.log 2>&1; tail -1 gpurun_out/ncu4_$L.log
done

# recogniser (SVTRv2-base): per-launch table, then one ncu --set full capture of the named plan steps
python tools/rec_bench.py --batch 1024 --dump gpurun_out/rec_launch_table.csv | tail -1
ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/rec_full -f \
  python tools/rec_bench.py --batch 1024 --profile-step stages.1.blocks.2.attn,stages.1.blocks.2.fc1,stages.1.blocks.2.fc2,stages.0.blocks.0.conv1,stages.1.blocks.2.norm1,stem.conv1 > gpurun_out/ncu_rec.log 2>&1
# hand-over kernel
ncu --set full --clock-control none --import-source on -k regex:rec_prep -s 3 -c 1 -o gpurun_out/recprep_full -f python tools/recprep_bench.py --iters 2 > gpurun_out/ncu_recprep.log 2>&1
