"""Shared pieces of the two sampler front ends: result files and the device-run wrapper."""
import os
import numpy as np


def write_chain_file(path, initmodel, obs, xmean, synmean, samples, syn):
    """Per-chain result file with the logical layout of the reference's HDF5 output
    (/root/reference/pyhmc/hmc.py:203-226,272-275): datasets `initmodel`, `obs`, groups
    `mean/{model,syn}` and `{i}/{model,syn}`.  h5py is not available in this image, so the same
    tree is stored in an .npz: keys "initmodel", "obs", "mean/model", "mean/syn", "models"
    ([nsamples,2n] == {i}/model stacked) and "syn" ([nsamples,ndata] == {i}/syn stacked).
    `tools/npz_to_h5.py` converts to the reference's exact HDF5 tree where h5py exists."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez(path, **{"initmodel": initmodel, "obs": obs, "mean/model": xmean, "mean/syn": synmean,
                      "models": samples, "syn": syn})


def best_mean_model(misfit, samples, nbest):
    """Average of the nbest lowest-misfit samples (hmc.py:266-270)."""
    idx = np.argsort(misfit)
    return np.mean(samples[idx[:nbest], :], axis=0)


def require_device_model(model):
    if not hasattr(model, "device_context"):
        raise TypeError(
            "rfsurfhmc_b200 samplers run the chains on the GPU and need a model exposing "
            "device_context(n) (rfsurfhmc_b200.model.model_rf_swd_vs_thk.Joint_RF_SWD); "
            "there is no CPU/Python fallback loop")
    return model
