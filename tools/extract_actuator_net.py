"""Extracts the 969 fp32 parameters of the reference's TorchScript actuator net
(resources/actuator_nets/anydrive_v3_lstm.pt, loaded at legged_gym/envs/anymal_c/anymal.py:54) into a
plain .npz the kernels and the oracle can load without TorchScript.  Run once in the build container."""
import numpy as np
import torch

SRC = "/root/reference/resources/actuator_nets/anydrive_v3_lstm.pt"
DST = "legged_gym_dev_b200/resources/anydrive_v3_lstm.npz"

m = torch.jit.load(SRC)
sd = {k: v.detach().numpy() for k, v in m.state_dict().items()}
out = {}
for l in (0, 1):
    out[f"w_ih{l}"] = sd[f"lstm.weight_ih_l{l}"]
    out[f"w_hh{l}"] = sd[f"lstm.weight_hh_l{l}"]
    out[f"b_ih{l}"] = sd[f"lstm.bias_ih_l{l}"]
    out[f"b_hh{l}"] = sd[f"lstm.bias_hh_l{l}"]
out["w_lin"], out["b_lin"] = sd["linear.weight"], sd["linear.bias"]
out["in_scale"] = sd["in_scale"].reshape(-1)
out["out_scale"] = sd["out_scale"].reshape(-1)
assert sum(v.size for k, v in out.items() if k[0] in "wb") == 969
np.savez(DST, **out)
print("wrote", DST, {k: v.shape for k, v in out.items()})
