"""Golden vectors for the policy/value network, made with the reference's shipped 3-block weights:

    python tests/golden/gen_net_golden.py      (build container only: needs /root/reference/models/model_5e-3_3_resnet.ot)

Writes tests/golden/net_golden.npz: the 60 tensors of the archive (VarStore names, '.' separators), 48 positions reached by random
play (planes [48,21,5,5]) and the outputs of the PyTorch twin of net.rs (onitama_alphazero_b200/net.py) on the CPU in f32:
policy [48,50], value [48]. The oracle's orc_net_forward and the tensor-core kernel are checked against these outputs
(tests/test_net_cpu.py, tests/test_gpu_parity.py). The weights are data of the reference repository, reproduced here only so that
the GPU box -- which has no /root/reference -- can run the comparison on a TRAINED network."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O
from onitama_alphazero_b200.net import ConvResNet

SRC = "/root/reference/models/model_5e-3_3_resnet.ot"


def main():
    model = ConvResNet(64, 21, 3).load_ot(SRC).eval()
    g = O.new_games(48, seed=2024)
    for s in range(48):                      # position i after (i mod 12) + 2 random plies
        live = np.arange(48) % 12 + 2 > s
        h = g.copy()
        O.env_step_random(h, 2024, s)
        g[live] = h[live]
    g = g[g["result"] == 0]
    planes = O.encode(g).reshape(-1, 21, 5, 5).astype(np.float32)
    with torch.no_grad():
        p, v = model(torch.from_numpy(planes))
    out = {"planes": planes, "policy": p.reshape(-1, 50).numpy(), "value": v.reshape(-1).numpy()}
    for k, t in model.state_dict().items():
        if not k.endswith("num_batches_tracked"):
            out["w:" + k] = t.detach().numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "net_golden.npz"), **out)
    print("positions", len(planes), "tensors", sum(1 for k in out if k.startswith("w:")), "policy max", float(out["policy"].max()),
          "value range", float(out["value"].min()), float(out["value"].max()))


if __name__ == "__main__":
    main()
