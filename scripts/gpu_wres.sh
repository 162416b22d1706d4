#!/bin/bash
# A/B of the resident-weight row-tiled ST-GCN GEMMs (gcn_w_res) + the GCN parity tests on the same box
mkdir -p gpurun_out
{
for i in 1 2; do
  python scripts/gcn_only.py 4096 gcn_w_res=0
  python scripts/gcn_only.py 4096 gcn_w_res=1
done
python scripts/gcn_only.py 512 gcn_w_res=0
python scripts/gcn_only.py 512 gcn_w_res=1
timeout 600 python -m pytest tests -m gpu -x -q -k "gcn or dropin or b4096 or smoke" 2>&1 | tail -5
} > gpurun_out/wres_ab.log 2>&1
cat gpurun_out/wres_ab.log
