"""Engine objects for the reference's own plug-in points.

``CvkEstimator`` is what ``ConditionalCFM.forward_estimator`` (cosyvoice/flow/flow_matching.py:126-153) dispatches to when
``self.estimator`` is not an ``nn.Module``: the reference then expects a TensorRT-style engine whose I/O contract is
``x[2,80,T] mask[2,1,T] mu[2,80,T] t[2] spks[2,80] cond[2,80,T]`` (contiguous, channel-major) with the result written over ``x``.
This class offers that contract as a plain callable on top of ``cvk_cfm_estimator_inplace`` so that a maintainer replaces the
TensorRT branch by ``x = self.estimator(x, mask, mu, t, spks, cond, streaming)``; the padding mask is implied by equal lengths
(the reference pads nothing inside one request: mask == 1 everywhere, flow/flow.py:262-263).
"""
import torch

from . import cvk


class CvkEstimator:
    def __init__(self, estimator_or_flow_state_dict, flow_cfg=(6, 4, 12, 4), precision="bf16", device=0, workspace_gb=4.0, context=None):
        """estimator_or_flow_state_dict: the state_dict of the reference's flow module (keys as in flow.pt) - the estimator's
        weights are the ``decoder.estimator.*`` entries; the encoder entries are needed only by the stage-level calls."""
        self.ctx = context or cvk.Context(device, precision, workspace_gb)
        if estimator_or_flow_state_dict is not None:
            self.ctx.load_state_dict("flow", estimator_or_flow_state_dict, list(flow_cfg))

    @torch.no_grad()
    def __call__(self, x, mask, mu, t, spks, cond, streaming=False):
        B, C, T = x.shape
        tm = lambda a: a.transpose(1, 2).reshape(B * T, C).contiguous().float()      # noqa: E731  [B,80,T] -> time-major [B*T,80]
        xt = tm(x).to(self.ctx.device)
        self.ctx.cfm_estimator_inplace(xt, tm(mu), t.float(), spks.float(), tm(cond), [T] * B, streaming=streaming)
        x.copy_(xt.view(B, T, C).transpose(1, 2).to(x.dtype))                         # the engine contract: result over x
        return x
