"""TEST INFRASTRUCTURE: fp32 restatement of the reference's learned label-embedding MLPs.

``model_y2h`` (CCDM_unified/models/resnet_y2h.py:143-173) and ``model_y2cov`` (models/resnet_y2cov.py:149-179) are the same
shape of network: ``y + 1e-8`` -> [Linear -> GroupNorm(8) -> ReLU] x 4 -> Linear -> ReLU, as called from
``LabelEmbed.fn_y2h`` / ``fn_y2cov`` (label_embedding.py:1028-1031, :1173-1176).  State-dict in (keys ``main.N.weight`` /
``main.N.bias`` of the reference's Sequential), tensor out.  Pinned to the reference's own outputs by
tests/test_label_mlp.py::test_oracle_vs_reference_golden (tests/golden/label_mlp.pt).
"""
import torch
import torch.nn.functional as F


def label_mlp_forward(sd, y, num_groups=8, eps=1e-5):
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    x = y.reshape(-1, 1).float() + 1e-8                     # resnet_y2h.py:170
    idx = sorted({int(k.split(".")[1]) for k in sd})
    i = 0
    while i < len(idx):
        n = idx[i]
        x = F.linear(x, sd[f"main.{n}.weight"], sd[f"main.{n}.bias"])
        if i + 1 < len(idx) and sd[f"main.{idx[i + 1]}.weight"].dim() == 1:          # GroupNorm affine follows this Linear
            g = idx[i + 1]
            x = F.group_norm(x, num_groups, sd[f"main.{g}.weight"], sd[f"main.{g}.bias"], eps)
            i += 1
        x = F.relu(x)
        i += 1
    return x
