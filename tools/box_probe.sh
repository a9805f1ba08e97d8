#!/bin/bash
# one-off environment probes on the GPU box
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv 2>&1 | head -4
nproc
python - <<'PY'
import os, sys, time, torch
sys.path.insert(0, '.')
from oracle import sg2_oracle as orc
from rewriting_b200.utils.stylegan2 import SeqStyleGAN2
from rewriting_b200.utils import zdataset
m = orc.seeded_state_dict(lambda: SeqStyleGAN2(256, 512, 8, mconv='seq'))
sd = dict(m.state_dict()); z = zdataset.standard_z_sample(2, 512, seed=1)
for th in (8, 16, 32, 64):
    torch.set_num_threads(th)
    with torch.no_grad():
        orc.generator_forward(sd, z)
        t0 = time.time(); orc.generator_forward(sd, z); dt = time.time() - t0
    print('threads', th, 'sec/fwd(B=2)', round(dt, 3), flush=True)
PY
