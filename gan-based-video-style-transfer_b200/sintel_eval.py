"""``computeTCL`` and the metric aggregation of the reference's Sintel evaluation.

Upstream: utils/sintel_eval.py:104-130 (``computeTCL``, ``save_dict_as_json``) and its arity variants
(ConGAN/CycleGAN/MoGAN ``sintel_eval.py:105``, StarGAN ``:105``, optimisation-based ``obst_eval.py:133``).
RAFT and the stylisation generator are upstream of the hot path (out of scope);
they are taken as callables.  What happens after them -- mask, warp, masked RMSE -- is one fused launch.
"""
import json
import os

import numpy as np
import torch

from . import ops


class InputPadder:
    """Pads images so that both dimensions are divisible by 8, like utils/raft/raft/utils/utils.py:7-24: replicate
    padding; mode 'sintel' splits the rows top/bottom (pad_ht // 2 on top), any other mode pads at the bottom only."""

    def __init__(self, dims, mode="sintel"):
        self.ht, self.wd = dims[-2:]
        pad_ht = (-self.ht) % 8
        pad_wd = (-self.wd) % 8
        left, top = pad_wd // 2, (pad_ht // 2 if mode == "sintel" else 0)
        self._pad = [left, pad_wd - left, top, pad_ht - top]

    def pad(self, *inputs):
        return [torch.nn.functional.pad(x, self._pad, mode="replicate") for x in inputs]

    def unpad(self, x):
        ht, wd = x.shape[-2:]
        return x[..., self._pad[2]:ht - self._pad[3], self._pad[0]:wd - self._pad[1]]


def computeRAFT(model, img1, img2, it=20, crop=True):
    """The reference's ``computeRAFT``: pad both images to multiples of 8 (``InputPadder(img1.shape)``, replicate mode,
    sintel split), run the flow estimator, return the full-resolution flow.  ``crop=True`` is the
    ConGAN/CycleGAN/MoGAN/StarGAN/fast_style_transfer/obst variant, ``flow_up[:, :, :H, :]`` (a view: rows from the TOP
    of the padded flow, not ``unpad`` -- the reference's own choice, ConGAN/sintel_eval.py:54-61); ``crop=False`` the
    StarGAN v2 one (utils/sintel_eval.py:53-60), which returns the padded flow as it is (its data set crops the frames to
    432 rows, so nothing was padded).  With H % 8 == 0 the two are the same.  The fused kernel reads the cropped view in place."""
    H = img1.shape[-2]
    with torch.no_grad():
        image1, image2 = InputPadder(img1.shape).pad(img1, img2)
        try:
            out = model(image1, image2, iters=it, test_mode=True)
        except TypeError:
            out = model(image1, image2)
    flow_up = out[-1] if isinstance(out, (tuple, list)) else out
    return flow_up[:, :, :H, :] if crop else flow_up


def computeTCL_from_flows(ff, bf, styled_prev, img_fake):
    """sqrt(mean((fbcCheckTorch(ff,bf) * (img_fake - warp(styled_prev, bf)))**2)) as a 0-dim tensor."""
    return ops.temporal_error(ff, bf, styled_prev, img_fake)


def _stylise(net, img, *extra):
    # the variants differ only in how the generator is invoked
    if hasattr(net, "generator"):
        return net.generator(img, *extra)
    if hasattr(net, "forward_eval") and not extra:
        return net.forward_eval(img)
    if hasattr(net, "run"):
        return net.run(img, *extra)
    return net(img, *extra)


def computeTCL(net, model, *args):
    """Drop-in for every ``computeTCL`` arity of the reference.

    * ``(net, model, s_trg, img_fake, img1, img2)``   utils/sintel_eval.py:104  (StarGAN v2: net.generator(img2, s_trg))
    * ``(net, model, img_fake, img1, img2)``          ConGAN/CycleGAN/MoGAN     (net.forward_eval(img2))
    * ``(net, model, img_fake, img1, img2, c_trg)``   StarGAN v1 (net(img2, c_trg)) and OBST (net.run(img2, sid))
    """
    if len(args) == 3:
        img_fake, img1, img2 = args
        extra = ()
    elif len(args) == 4:
        last_is_image = (torch.is_tensor(args[3]) and args[3].dim() == 4 and torch.is_tensor(args[2])
                         and args[3].shape == args[2].shape)
        if last_is_image:   # (s_trg, img_fake, img1, img2)
            extra, img_fake, img1, img2 = (args[0],), args[1], args[2], args[3]
        else:               # (img_fake, img1, img2, c_trg | sid)
            img_fake, img1, img2, extra = args[0], args[1], args[2], (args[3],)
    else:
        raise TypeError("computeTCL expects (net, model, [s_trg,] img_fake, img1, img2[, c_trg])")
    ff_last = computeRAFT(model, img2, img1)
    bf_last = computeRAFT(model, img1, img2)
    with torch.no_grad():
        styled_prev = _stylise(net, img2, *extra)
    return computeTCL_from_flows(ff_last, bf_last, styled_prev, img_fake)


def aggregate_means(out_id, data_dict, num_domains):
    """The arithmetic of ``save_dict_as_json`` (utils/sintel_eval.py:112-126) without the file write."""
    data_dict = dict(data_dict)
    n = len(data_dict)
    dict_mean = 0
    dict_mean_s = np.zeros(num_domains - 1)
    for key, value in data_dict.items():
        dict_mean += value / n
        for d in range(1, num_domains):
            if ("_s" + str(d)) in key:
                dict_mean_s[d - 1] += value / (n / 3)
    data_dict[out_id + "_mean"] = float(dict_mean)
    for d in range(1, num_domains):
        data_dict[out_id + "_mean_s" + str(d)] = float(dict_mean_s[d - 1])
    return data_dict


def save_dict_as_json(out_id, data_dict, out_path, num_domains):
    out = aggregate_means(out_id, data_dict, num_domains)
    data_dict.update(out)
    with open(os.path.join(out_path, out_id + ".json"), "w") as f:
        json.dump(data_dict, f, indent=4, sort_keys=False)
