"""Batched device API: torch tensors in, torch tensors out (PyTorch only hands over device memory
and the current stream; all numerics are in librfsurf_b200.so)."""
import torch
from ._lib import Context


class JointEvaluator:
    """Many-chains Joint_RF_SWD.misfit_and_grad on one GPU with resident inputs/outputs."""

    def __init__(self, cfg, dobs, nlayer, device=0, which=0):
        self.ctx = Context(device)
        self.which = which
        self.n = nlayer
        if which != 1:
            self.ctx.config_swd(nlayer, cfg.get("tRc"), cfg.get("tRg"), cfg.get("tLc"), cfg.get("tLg"),
                                mode=cfg.get("mode", 0), sphere=cfg.get("sphere", False),
                                stale=cfg.get("stale", True))
        if which != 2:
            self.ctx.config_rf(nlayer, cfg["ray_p"], cfg["nt"], cfg["dt"], cfg["gauss"], cfg["time_shift"],
                               cfg.get("water", 0.001), cfg.get("rf_type", "P"), cfg.get("method", "freq"))
        self.ctx.config_obs(dobs, cfg.get("sigma1", 1.0), cfg.get("sigma2", 1.0))
        self.device = torch.device("cuda", device)
        self.ndata = self.ctx.ndata(which)
        self._out = None

    def __call__(self, x):
        """x: float64 CUDA tensor [B, 2n] -> (U [B], grad [B,2n], dsyn [B,ndata], flag [B] uint8)."""
        assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous()
        B = x.shape[0]
        if self._out is None or self._out[0].shape[0] != B:
            self._out = (torch.empty(B, dtype=torch.float64, device=self.device),
                         torch.empty(B, 2 * self.n, dtype=torch.float64, device=self.device),
                         torch.empty(B, self.ndata, dtype=torch.float64, device=self.device),
                         torch.empty(B, dtype=torch.uint8, device=self.device))
        U, G, D, F = self._out
        self.ctx.misfit_grad_dev(B, x.data_ptr(), self.which, U.data_ptr(), G.data_ptr(), D.data_ptr(),
                                 F.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
        return U, G, D, F
