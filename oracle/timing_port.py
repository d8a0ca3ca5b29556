"""ORACLE support -- test/bench infrastructure, not product code.

The reference's op sequence for the slice-to-3D path rebuilt from a flat state_dict with the
SAME PyTorch library operators the reference module calls (F.conv2d / F.batch_norm,
F.grid_sample, nn.TransformerEncoder's fused eval fast path, nn.Linear), so that the CPU
baseline timed by bench.py runs at the speed of the reference itself rather than at the
speed of oracle/oracle.py's explicit restatement (which spells grid_sample and attention out
with index arithmetic and is ~2x slower).  Values agree with oracle.py to ~1e-6
(tests/test_oracle_golden.py::test_timing_port_matches_oracle).

The reference package itself cannot travel to the GPU box (it needs /root/reference, which
does not exist there), hence this port; cited lines are relative to reg_slices/.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import oracle


class TimingPort:
    def __init__(self, sd, n_slices=12):
        self.sd, self.K = sd, n_slices
        layer = nn.TransformerEncoderLayer(d_model=128, nhead=4, batch_first=True)  # src/models.py:18
        self.att = nn.TransformerEncoder(layer, num_layers=3)                       # src/models.py:19
        self.att.load_state_dict({k[len("att_decoder."):]: v for k, v in sd.items() if k.startswith("att_decoder.")})
        self.att.eval()

    @torch.no_grad()
    def decode(self, feats, qry, T):
        """src/models.py:68-84 with the reference's own operators.  qry (1,M,3) already flipped."""
        sd, K = self.sd, self.K
        B, M, _ = qry.shape
        uv = oracle.project_coord(qry, T)
        grid = uv.view(B, 1, M, 2).expand(-1, K, -1, -1).reshape(B * K, 1, M, 2)
        samp = [F.grid_sample(f, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
                .permute(0, 3, 2, 1).reshape(B * K, M, f.shape[1]) for f in feats]
        agg = torch.cat(samp, 2).view(B, K, M, 992).permute(0, 2, 1, 3).reshape(B * M, K, 992)
        tok = torch.cat([F.linear(qry, sd["fc_p.weight"], sd["fc_p.bias"]).view(B * M, 1, 128),
                         F.linear(agg, sd["fc_s.weight"], sd["fc_s.bias"])], 1)
        x = self.att(tok).view(B, M, K + 1, 128)[:, :, 0, :]
        return F.linear(x, sd["fc_out.0.weight"], sd["fc_out.0.bias"]).squeeze(-1)

    @torch.no_grad()
    def forward_as_written(self, feed):
        """One Slices3DRegModel.forward in test mode (src/models.py:48-94): U-Net, decoder and the
        VGG19 perceptual loss the reference evaluates (and discards) on every inference chunk."""
        sd, K = self.sd, self.K
        q = oracle.prepare_queries(feed["qry_norot"], None, "test")
        feats, rec = oracle.unet_forward(sd, feed["img_input"], K)
        sdf = self.decode(feats, q, feed["trans_mat_wo_rot_tp"])
        B, _, S, _ = feed["img_input"].shape
        vgg = oracle.vgg_perceptual(sd, rec, feed["img_slices"].view(B * K, 3, S, S)) * 0.001
        return sdf, vgg
