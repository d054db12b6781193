#!/bin/bash
# end-of-round evidence run (round 2), part B: ncu --set full of the kernels of this round; the report stays on the box
# (it is larger than what travels back), its raw page is exported as CSV
set -x
L=smooth0,layer3.1.dcn,layer2.1.dcn,layer4.1.dcn,asf,head.bin.1,head.0,layer1.0.conv3,layer1.1.conv3,layer1.2.conv3,layer2.0.conv3,layer3.0.conv3,layer4.0.conv3,layer3.1.conv1,layer3.1.conv3,layer2.1.offset,layer1.1.conv2,maxpool,stem
P="python bench.py --profile-launch $L"
$P > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --nvtx --nvtx-include "regex:profile:.*/" -f -o /tmp/r02_ncu_full $P > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-200
ncu -i /tmp/r02_ncu_full.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_raw.csv 2>/dev/null; ls -la gpurun_out/
