"""Golden vectors for get_rope_index: outputs of the reference's own
InfiniteVLModel.get_rope_index (infinitevl_standard/modeling_infinitevl.py:1623-1758) on synthetic token rows with
image and video placeholders.  Run in the BUILD container only (needs /root/reference):
    python tests/golden/make_golden_rope_index.py   ->  tests/golden/ref_rope_index.npz"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def cases():
    IMG, VID, VS, VE = 151655, 151656, 151652, 151653
    def block(kind, t, h, w):
        return [VS] + [kind] * (t * (h // 2) * (w // 2)) + [VE]
    out = []
    # 1: text, image 1x4x6, text
    ids = list(range(10, 17)) + block(IMG, 1, 4, 6) + list(range(20, 25))
    out.append(dict(ids=[ids], img=[[1, 4, 6]], vid=None, spg=None, mask=None))
    # 2: two images and a video (2 temporal grids), seconds per grid 1.5, batch of 2 with left padding
    a = [5, 6] + block(IMG, 1, 8, 8) + [7] + block(VID, 2, 4, 4) + [8, 9, 10] + block(IMG, 1, 2, 4) + [11]
    b = [3] * 9 + block(IMG, 1, 4, 4) + [12, 13]
    L = max(len(a), len(b))
    pad = lambda r: [0] * (L - len(r)) + r
    mask = [[0] * (L - len(a)) + [1] * len(a), [0] * (L - len(b)) + [1] * len(b)]
    out.append(dict(ids=[pad(a), pad(b)], img=[[1, 8, 8], [1, 2, 4], [1, 4, 4]], vid=[[2, 4, 4]], spg=[1.5], mask=mask))
    # 3: stream frames: 3 frames of 16x16 patches (8x8 tokens) each followed by text (SURVEY.md 8d generator)
    ids = list(range(30, 46))
    for f in range(3):
        ids += block(IMG, 1, 16, 16) + list(range(100, 120))
    out.append(dict(ids=[ids], img=[[1, 16, 16]] * 3, vid=None, spg=None, mask=None))
    # 4: video only, default seconds per grid
    ids = [1, 2] + block(VID, 3, 4, 8) + [3]
    out.append(dict(ids=[ids], img=None, vid=[[3, 4, 8]], spg=None, mask=None))
    return out


def main():
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS
    ROPE_INIT_FUNCTIONS.setdefault("default", lambda config, device=None, **kw: (torch.ones(64), 1.0))
    sys.path.insert(0, os.path.join(REF, "infinitevl"))
    import infinitevl_standard.modeling_infinitevl as M
    cfg = types.SimpleNamespace(vision_config=types.SimpleNamespace(spatial_merge_size=2, tokens_per_second=2),
                                image_token_id=151655, video_token_id=151656, vision_start_token_id=151652)
    fake = types.SimpleNamespace(config=cfg)
    save = {}
    for i, c in enumerate(cases()):
        t = lambda x, dt=torch.long: None if x is None else torch.tensor(x, dtype=dt)
        pos, delta = M.InfiniteVLModel.get_rope_index(fake, t(c["ids"]), t(c["img"]), t(c["vid"]),
                                                      second_per_grid_ts=t(c["spg"], torch.float32),
                                                      attention_mask=t(c["mask"]))
        save[f"pos{i}"] = pos.numpy()
        save[f"delta{i}"] = delta.numpy()
    np.savez_compressed(os.path.join(HERE, "ref_rope_index.npz"), **save)
    print("wrote ref_rope_index.npz", {k: v.shape for k, v in save.items()})


if __name__ == "__main__":
    main()
