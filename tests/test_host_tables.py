"""CPU checks of the host-side tables behind the one-launch-per-step kernels (gradient fold, weight repack): every
element of every parameter is covered by exactly one (block, offset) pair. No kernel is launched."""
import numpy as np
import torch


def _cover(table, numels):
    chunk_of = {}
    seen = [np.zeros(n, dtype=np.int32) for n in numels]
    key_item = "blk_seg" if "blk_seg" in table else "blk_item"
    items, starts = table[key_item].tolist(), table["blk_start"].tolist()
    return items, starts, seen


def test_grad_fold_table_covers_every_element_once():
    from b200seg import raw, _lib
    chunk = _lib.lib().b200seg_grad_fold_chunk()
    segs = [(0, 0, 48, 32, 9, 32), (13824, 13824, 1, 96, 1, 96), (13952, 13952, 19, 512, 1, 512),
            (23680, 23680, 96, 96, 9, 96), (106624, 200000, 64, 3, 9, 16)]
    tab = raw.grad_fold_table(segs, "cpu")
    numels = [s[2] * s[3] * s[4] for s in segs]
    items, starts, seen = _cover(tab, numels)
    assert tab["n_blocks"] == len(items) == sum((n + chunk - 1) // chunk for n in numels)
    for it, st in zip(items, starts):
        seen[it][st:st + chunk] += 1
    assert all((s == 1).all() for s in seen)
    raw_segs = np.frombuffer(tab["segs"].numpy().tobytes(), dtype=np.dtype(
        [("offset", "<i8"), ("src_offset", "<i8"), ("cout", "<i4"), ("cin", "<i4"), ("taps", "<i4"), ("src_cin", "<i4")]))
    assert [tuple(int(v) for v in r) for r in raw_segs] == segs          # 32-byte records, C layout


def test_pack_table_records_pointers_and_blocks():
    from b200seg import raw, _lib
    chunk = _lib.lib().b200seg_pack_chunk()
    w0 = torch.randn(64, 3, 3, 3)
    w1 = torch.randn(19, 512, 1, 1)
    f0, d0 = torch.zeros(64, 9, 16, dtype=torch.bfloat16), torch.zeros(16, 9, 64, dtype=torch.bfloat16)
    f1, d1 = torch.zeros(19, 1, 512, dtype=torch.bfloat16), torch.zeros(512, 1, 24, dtype=torch.bfloat16)
    tab = raw.pack_table([(w0, f0, d0), (w1, f1, d1)], "cpu")
    rec = np.frombuffer(tab["items"].numpy().tobytes(), dtype=np.dtype(
        [("w", "<u8"), ("f", "<u8"), ("d", "<u8"), ("o", "<i4"), ("i", "<i4"), ("i_dst", "<i4"), ("k", "<i4"),
         ("o_pad", "<i4"), ("r", "<i4")]))
    assert int(rec[0]["w"]) == w0.data_ptr() and int(rec[1]["d"]) == d1.data_ptr()
    assert (int(rec[0]["i"]), int(rec[0]["i_dst"]), int(rec[0]["o_pad"])) == (3, 16, 64)
    assert (int(rec[1]["o"]), int(rec[1]["k"]), int(rec[1]["o_pad"])) == (19, 1, 24)
    assert tab["n_blocks"] == (w0.numel() + chunk - 1) // chunk + (w1.numel() + chunk - 1) // chunk
    assert tab["ptrs"] == (w0.data_ptr(), w1.data_ptr())


def test_iou_from_hist():
    from b200seg.evaltail import iou_from_hist
    h = torch.tensor([[5, 1], [2, 7]])
    iou = iou_from_hist(h)
    assert torch.allclose(iou, torch.tensor([5 / 8, 7 / 10], dtype=torch.float64))


def test_syncbn_exchange_layout_regions_are_disjoint():
    from b200seg.p2p import exchange_layout
    chans = {"a.bn1": 48, "a.bn2": 96, "b.bn": 720, "c.bn": 19}
    world, passes = 4, 2
    mail, flag, msize, fsize = exchange_layout(chans, world, passes)
    assert len(mail) == len(flag) == passes * len(chans) * 2
    spans = sorted((mail[k], mail[k] + world * 2 * chans[k[1]]) for k in mail)
    assert spans[0][0] == 0 and spans[-1][1] == msize
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))              # contiguous, non-overlapping
    fsp = sorted((flag[k], flag[k] + world * ((chans[k[1]] + 31) // 32)) for k in flag)
    assert fsp[0][0] == 0 and fsp[-1][1] == fsize and all(a[1] == b[0] for a, b in zip(fsp, fsp[1:]))
    assert mail[(0, "a.bn1", 1)] == world * 2 * 48                           # direction 1 follows direction 0
