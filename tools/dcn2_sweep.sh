#!/bin/bash
# development sweep of the tensor-memory DCN kernel's plan knobs on the stage shapes (batch 16)
for tw in 16 32; do for bs in 2 3 4; do for eb in 1 2; do
  echo "== tw $tw bstages $bs epibufs $eb"
  B200_CONV_TW=$tw B200_DCN2_BSTAGES=$bs B200_DCN2_EPIBUFS=$eb timeout 100 python tools/dcn_microbench.py --batches 16 --only "s1" --no-tv --out gpurun_out/x.csv 2>&1 | grep -E "layer2|layer3" | cut -d, -f1,8,10,11
done; done; done
