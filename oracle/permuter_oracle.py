"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the dual-grain code permuter of the stage-2 tokenisation
path: a numpy restatement of ``/root/reference/modules/dynamic_modules/permuter.py:50-132``.
Only ``tests/`` may import this module - it is the checker, never the product path.

Pinning: the reference's only "test" of this class is the round-trip print in its ``__main__`` block
(:136-306, forward then forward_back reproduces the code map); the oracle is pinned against outputs
of the reference class itself run in the build container (``tests/golden/make_golden.py`` ->
``tests/golden/permuter.npz``; checked by ``tests/test_oracle_golden.py``), including that round trip.
"""
import numpy as np

DEFAULT_CODES = dict(content_pad_code=1024, content_eos_code=1025, coarse_position_pad_code=256,
                     coarse_position_eos_code=257, fine_position_pad_code=1024, fine_position_eos_code=1025)


def _pad(seqs, pad):
    """torch.nn.utils.rnn.pad_sequence(batch_first=True): right-pad to the longest sequence."""
    n = max(len(s) for s in seqs)
    out = np.full((len(seqs), n), pad, np.int64)
    for i, s in enumerate(seqs):
        out[i, :len(s)] = s
    return out


def forward(indices, grain, coarse_hw=16, fine_hw=32, fine_position_order="region-first", **codes):
    """permuter.py:50-109.  indices [B,F,F] int64, grain [B,Hc,Hc] (0 coarse, 1 fine)."""
    c = dict(DEFAULT_CODES, **codes)
    hw1, hw2 = coarse_hw, fine_hw // coarse_hw
    B = indices.shape[0]
    # "B (h1 h2) (w1 w2) -> B h1 w1 (h2 w2)" (:56)
    reg = indices.reshape(B, hw1, hw2, hw1, hw2).transpose(0, 1, 3, 2, 4).reshape(B, hw1, hw1, hw2 * hw2)
    pos_coarse = np.arange(hw1 * hw1)
    pos_fine = np.arange(fine_hw * fine_hw).reshape(fine_hw, fine_hw)
    cc, cp, fc, fp = [], [], [], []
    for i in range(B):
        m0 = grain[i] == 0
        cc.append(np.concatenate([reg[i, :, :, 0][m0], [c["content_eos_code"]]]))                  # :60-61
        cp.append(np.concatenate([pos_coarse[m0.reshape(-1)], [c["coarse_position_eos_code"]]]))    # :71-72
        if fine_position_order == "region-first":
            m1 = grain[i] == 1
            pf = pos_fine.reshape(hw1, hw2, hw1, hw2).transpose(0, 2, 1, 3).reshape(hw1, hw1, hw2 * hw2)
            fc.append(np.concatenate([reg[i][m1].reshape(-1), [c["content_eos_code"]]]))            # :80-81
            fp.append(np.concatenate([pf[m1].reshape(-1), [c["fine_position_eos_code"]]]))          # :84-85
        else:
            m1 = np.repeat(np.repeat(grain[i], hw2, axis=-1), hw2, axis=-2) == 1                    # :88
            fc.append(np.concatenate([indices[i][m1].reshape(-1), [c["content_eos_code"]]]))        # :89-90
            fp.append(np.concatenate([pos_fine[m1], [c["fine_position_eos_code"]]]))                # :93-94
    out = {
        "coarse_content": _pad(cc, c["content_pad_code"]),
        "fine_content": _pad(fc, c["content_pad_code"]),
        "coarse_position": _pad(cp, c["coarse_position_pad_code"]),
        "fine_position": _pad(fp, c["fine_position_pad_code"]),
    }
    out["coarse_segment"] = np.zeros_like(out["coarse_content"])                                   # :75
    out["fine_segment"] = np.ones_like(out["fine_content"])                                        # :99
    return out


def forward_back(coarse_content, fine_content, coarse_position, fine_position, coarse_hw=16, fine_hw=32, **codes):
    """permuter.py:111-132, element by element like the reference (later elements overwrite earlier
    ones; the coarse map is spread only once its eos is met)."""
    c = dict(DEFAULT_CODES, **codes)
    hw1, hw2 = coarse_hw, fine_hw // coarse_hw
    B = coarse_content.shape[0]
    target = np.zeros((B, fine_hw * fine_hw), np.int64)
    for i in range(B):
        coarse = np.zeros(hw1 * hw1, np.int64)
        for j in range(coarse_content.shape[1]):
            if coarse_position[i, j] == c["coarse_position_eos_code"]:
                # repeat_interleave + "(h1 w1 h2 w2) -> (h1 h2 w1 w2)" (:118-119): every fine slot of a cell
                target[i] = np.repeat(np.repeat(coarse.reshape(hw1, hw1), hw2, 0), hw2, 1).reshape(-1)
                break
            coarse[coarse_position[i, j]] = coarse_content[i, j]
        for j in range(fine_content.shape[1]):
            if fine_position[i, j] == c["fine_position_eos_code"]:
                break
            target[i, fine_position[i, j]] = fine_content[i, j]
    return target.reshape(B, fine_hw, fine_hw)
