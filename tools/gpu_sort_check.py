#!/usr/bin/env python
"""Standalone check of gsr_sort_pairs (onesweep) against torch.sort(stable=True)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_deformable_b200"))
import gsr_runtime as rt  # noqa: E402


def sort_pairs(keys, vals, begin_bit, end_bit):
    lib = rt.load()
    n = keys.numel()
    ka, kb = keys.clone(), torch.zeros_like(keys)
    va, vb = vals.clone(), torch.zeros_like(vals)
    nbytes = lib.gsr_sort_bytes(n, begin_bit, end_bit)
    temp = torch.zeros(max(nbytes, 1), dtype=torch.uint8, device=keys.device)
    in_b = ctypes.c_int(0)
    rt.check(lib.gsr_sort_pairs(rt.ptr(ka), rt.ptr(kb), rt.ptr(va), rt.ptr(vb), n, begin_bit, end_bit,
                                rt.ptr(temp), nbytes, ctypes.byref(in_b), rt.stream_ptr()))
    torch.cuda.synchronize()
    err = int(temp[(8 * 256 + 63) * 4:(8 * 256 + 64) * 4].view(torch.int32).item()) if n else 0
    return (kb, vb) if in_b.value else (ka, va), err


def main():
    g = torch.Generator().manual_seed(0)
    ok = True
    for n, bits, kind in ((0, 45, "rand"), (1, 45, "rand"), (5, 45, "rand"), (4096, 45, "rand"), (4097, 45, "rand"),
                          (100003, 45, "rand"), (100003, 45, "dups"), (1 << 20, 47, "rand"), (3000017, 44, "tile"),
                          (8000000, 45, "tile"), (1 << 20, 13, "rand"), (1 << 20, 64, "rand")):
        if kind == "rand":
            keys = torch.randint(0, 2 ** 62, (n,), generator=g, dtype=torch.int64)
        elif kind == "dups":
            keys = torch.randint(0, 50, (n,), generator=g, dtype=torch.int64) << 20
        else:  # tile|depth like keys
            tile = torch.randint(0, 8160, (n,), generator=g, dtype=torch.int64)
            depth = (torch.rand((n,), generator=g) * 5 + 0.2).view(torch.int32).to(torch.int64)
            keys = (tile << 32) | depth
        if bits < 64:
            keys_full = keys | (torch.randint(0, 3, (n,), generator=g, dtype=torch.int64) << 62 if False else 0)
        keys = keys.cuda()
        vals = torch.arange(n, dtype=torch.int32).cuda()
        (ks, vs), err = sort_pairs(keys, vals, 0, bits)
        mask = (1 << bits) - 1 if bits < 63 else -1
        if bits >= 63:
            ref_keys = keys.clone()
            # unsigned order for 64-bit: flip sign bit
            order = torch.sort(keys ^ (-2 ** 63), stable=True).indices if bits == 64 else torch.sort(keys, stable=True).indices
        else:
            order = torch.sort(keys & mask, stable=True).indices
        good = bool(torch.equal(ks, keys[order])) and bool(torch.equal(vs.long(), order)) and err == 0
        ok &= good
        print("sort n=%d bits=%d kind=%s: %s (err flag %d)" % (n, bits, kind, "OK" if good else "MISMATCH", err))
        if n >= 1 << 20:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            lib = rt.load()
            ka, kb, va, vb = keys.clone(), torch.zeros_like(keys), vals.clone(), torch.zeros_like(vals)
            nbytes = lib.gsr_sort_bytes(n, 0, bits)
            temp = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
            in_b = ctypes.c_int(0)
            ts = []
            for it in range(5):
                ka.copy_(keys); va.copy_(vals)
                e0.record()
                lib.gsr_sort_pairs(rt.ptr(ka), rt.ptr(kb), rt.ptr(va), rt.ptr(vb), n, 0, bits, rt.ptr(temp), nbytes,
                                   ctypes.byref(in_b), rt.stream_ptr())
                e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = sorted(ts)[len(ts) // 2]
            passes = (bits + 7) // 8
            print("   time %.3f ms  -> %.0f GB/s (8+24*passes B/pair model)" % (t, n * (8 + 24 * passes) / t / 1e6))
            ts = []
            for it in range(5):
                e0.record(); torch.sort(keys); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            print("   torch.sort(keys only, 64-bit) %.3f ms" % sorted(ts)[2])
    print("SORT_ALL_OK" if ok else "SORT_FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
