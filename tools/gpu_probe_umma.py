"""Pins down tcgen05 shared-memory descriptor semantics on real hardware (run under gpurun).

Questions answered (each prints MATCH / MISMATCH plus a diagnosis):
  Q1  sanity: K-major SWIZZLE_128B operands written by TMA, D = A * I.
  Q2  can the A descriptor start at a row offset that is not a multiple of the 1024-byte swizzle repeat
      (shifted im2col windows over one halo tile)? With base_offset 0 / base_offset = row & 7.
  Q3  stride-byte-offset other than 1024 (8-row groups at a pitch of 16 or 10 rows).
  Q4  MN-major operands (weight-gradient GEMM: both operands are [pixels][channels]).
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))
import torch

from b200seg._lib import ProbeDesc, ptr, stream_ptr, test_lib


def operand(o, rows, cols, box_cols, box_rows, nboxes=1, c0=0, r0=0, dcol=0, drow=0, smem_stride=0, swz=128):
    o.rows, o.cols, o.box_cols, o.box_rows = rows, cols, box_cols, box_rows
    o.nboxes, o.c0, o.r0, o.dcol, o.drow, o.smem_stride, o.swizzle_bytes = nboxes, c0, r0, dcol, drow, smem_stride, swz


def run(p, A, B, N):
    D = torch.full((128, N), float("nan"), device="cuda")
    rc = test_lib().b200seg_umma_probe(ctypes.byref(p), ptr(A), ptr(B), ptr(D), stream_ptr())
    assert rc == 0, rc
    torch.cuda.synchronize()
    return D


def diagnose(D, A):
    """For the first rows of D report which source row of A (and 16-byte chunk permutation) they equal."""
    out = []
    Af = A.float()
    for m in list(range(0, 10)) + [16, 17, 127]:
        row = D[m]
        hit = None
        for r in range(Af.shape[0]):
            if torch.equal(row, Af[r]):
                hit = "row %d" % r
                break
        if hit is None:
            # try chunk permutations: match each 8-element chunk separately
            chunks = []
            for c in range(8):
                seg = row[c * 8:(c + 1) * 8]
                found = "?"
                for r in range(Af.shape[0]):
                    for c2 in range(8):
                        if torch.equal(seg, Af[r, c2 * 8:(c2 + 1) * 8]):
                            found = "%d.%d" % (r, c2)
                            break
                    if found != "?":
                        break
                chunks.append(found)
            hit = "chunks " + ",".join(chunks)
        out.append("m%d<-%s" % (m, hit))
    return " ".join(out)


def main():
    g = torch.Generator().manual_seed(0)
    # distinct-ish small integers, exact in bf16
    A = torch.randint(-100, 100, (256, 64), generator=g).float().to(torch.bfloat16).cuda()
    I = torch.eye(64).to(torch.bfloat16).cuda()

    def base_desc():
        p = ProbeDesc()
        operand(p.a, 256, 64, 64, 256)
        operand(p.b, 64, 64, 64, 64)
        p.M, p.N, p.ksteps = 128, 64, 4
        p.a_off, p.a_lbo, p.a_sbo, p.a_layout, p.a_base, p.a_major, p.a_kstep = 0, 16, 1024, 2, 0, 0, 32
        p.b_off, p.b_lbo, p.b_sbo, p.b_layout, p.b_base, p.b_major, p.b_kstep = 0, 16, 1024, 2, 0, 0, 32
        return p

    p = base_desc()
    D = run(p, A, I, 64)
    print("Q1 sanity K-major SW128:", "MATCH" if torch.equal(D, A[:128].float()) else "MISMATCH " + diagnose(D, A),
          flush=True)

    for shift in (1, 2, 3, 5, 8, 9):
        for base in sorted({0, shift & 7}):
            p = base_desc()
            p.a_off, p.a_base = shift * 128, base
            D = run(p, A, I, 64)
            exp = A[shift:shift + 128].float()
            print("Q2 shift %d rows base_offset %d:" % (shift, base),
                  "MATCH" if torch.equal(D, exp) else "MISMATCH " + diagnose(D, A), flush=True)

    for pitch_rows in (16, 10, 12):
        for shift in (0, 1, 2):
            p = base_desc()
            p.a_sbo = pitch_rows * 128
            p.a_off = shift * 128
            D = run(p, A, I, 64)
            idx = torch.tensor([(m // 8) * pitch_rows + (m % 8) + shift for m in range(128)])
            ok = idx.max().item() < 256 and torch.equal(D, A[idx.cuda()].float())
            print("Q3 sbo=%d rows shift %d:" % (pitch_rows, shift), "MATCH" if ok else "MISMATCH " + diagnose(D, A),
                  flush=True)

    # Q4: MN-major. At [K=64][M=128], Bt [K=64][N=64]; D = At^T @ Bt
    At = torch.randint(-4, 5, (64, 128), generator=g).float().to(torch.bfloat16).cuda()
    Bt = torch.randint(-4, 5, (64, 64), generator=g).float().to(torch.bfloat16).cuda()
    exp = At.float().t() @ Bt.float()
    for (lbo, sbo, tag) in ((8192, 1024, "lbo=chunk sbo=8rows"), (1024, 8192, "swapped")):
        p = ProbeDesc()
        operand(p.a, 64, 128, 64, 64, nboxes=2, dcol=64, smem_stride=8192)
        operand(p.b, 64, 64, 64, 64)
        p.M, p.N, p.ksteps = 128, 64, 4
        p.a_off, p.a_lbo, p.a_sbo, p.a_layout, p.a_base, p.a_major, p.a_kstep = 0, lbo, sbo, 2, 0, 1, 2048
        p.b_off, p.b_lbo, p.b_sbo, p.b_layout, p.b_base, p.b_major, p.b_kstep = 0, lbo, sbo, 2, 0, 1, 2048
        D = run(p, At, Bt, 64)
        ok = torch.equal(D, exp)
        print("Q4 MN-major SW128 (%s):" % tag, "MATCH" if ok else "MISMATCH maxdiff %.1f" % (D - exp).abs().max().item(),
              flush=True)
    # Q4b: MN-major A with K-major B (mixed)
    Bk = Bt.t().contiguous()   # [N=64][K=64] K-major
    p = ProbeDesc()
    operand(p.a, 64, 128, 64, 64, nboxes=2, dcol=64, smem_stride=8192)
    operand(p.b, 64, 64, 64, 64)
    p.M, p.N, p.ksteps = 128, 64, 4
    p.a_off, p.a_lbo, p.a_sbo, p.a_layout, p.a_base, p.a_major, p.a_kstep = 0, 8192, 1024, 2, 0, 1, 2048
    p.b_off, p.b_lbo, p.b_sbo, p.b_layout, p.b_base, p.b_major, p.b_kstep = 0, 16, 1024, 2, 0, 0, 32
    D = run(p, At, Bk, 64)
    print("Q4b MN-major A x K-major B:", "MATCH" if torch.equal(D, exp) else "MISMATCH", flush=True)


if __name__ == "__main__":
    main()
