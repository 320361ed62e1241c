"""Benchmark workloads (BASELINE.json configs) as seeded synthetic inputs, generated on the device.

``configs[1]``: 1,000 synthetic 5 Mb Klebsiella-like assemblies vs a kpsc_k-shaped database
(150 loci x 20 genes, 4 core gene families shared by all loci) on one B200.
``configs[2]`` (the configuration the metric is quoted on): 10,000 assemblies vs kpsc_k + kpsc_o in ONE index
(``synth.make_ko_db``), a K locus and an O locus embedded in every assembly (``locus_ranges``).

The 5 Gbase of background sequence are drawn with torch on the GPU (numpy would take minutes);
the embedded locus of each assembly is mutated on the host (small) and copied in.  torch is used
for device memory and RNG only.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import synth


@dataclass
class DeviceWorkload:
    ascii: "object"  # torch.uint8 tensor [n_asm * asm_len] on the device
    contig_off: np.ndarray  # int64
    contig_len: np.ndarray  # int32
    asm_contig_start: np.ndarray  # int32, n_asm + 1
    locus: np.ndarray  # embedded loci per assembly, [n_asm, n_classes] (K locus, O locus, ...)
    n_asm: int
    asm_len: int

    def host_assembly(self, a: int) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Assembly `a` as host arrays (ASCII bytes, int64 contig offsets, int32 lengths): what the CPU legs of bench.py map, so
        that they see exactly the sequences the GPU step mapped."""
        seq = self.ascii[a * self.asm_len : (a + 1) * self.asm_len].cpu().numpy()
        c0, c1 = int(self.asm_contig_start[a]), int(self.asm_contig_start[a + 1])
        return seq, (self.contig_off[c0:c1] - a * self.asm_len).astype(np.int64), self.contig_len[c0:c1].astype(np.int32)


def make_device_workload(db: synth.SynthDB, n_asm: int, asm_len: int = 5_000_000, mean_contigs: float = 80.0,
                         seed: int = 1000, device: str = "cuda:0", gc: float = 0.57, n_frac: float = 1e-4,
                         sub: tuple[float, float] = (0.0, 0.05), indel: tuple[float, float] = (0.0, 0.005),
                         first_index: int = 0, locus_ranges: tuple[tuple[int, int], ...] | None = None, repeats: int = 0,
                         repeat_seq: bytes | None = None) -> DeviceWorkload:
    """Assembly `first_index + a` is a pure function of (seed, first_index + a, asm_len): the background is drawn in fixed chunks
    of `step` assemblies whose generator seed depends on the chunk's first global index only, so a rank, a smaller sample (the CPU
    arm) and the full batch all see the same sequences for the same global assembly index.

    repeats > 0 writes that many copies of `repeat_seq` (an insertion-sequence-like element; either strand) into every assembly,
    away from the embedded loci: when the element is also a database gene, its minimizers occur more often than minimap2's
    min_mid_occ floor and every assembly takes the occurrence-census path (mm_idx_cal_max_occ)."""
    import torch

    dev = torch.device(device)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    out = torch.empty(n_asm * asm_len, dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev)
    thr = torch.tensor([(1 - gc) / 2, 0.5, 0.5 + gc / 2], device=dev)  # A | C | G | T cumulative
    step = max(1, (256 << 20) // asm_len)
    # chunks are aligned to multiples of `step` in GLOBAL assembly indices and always drawn at full size
    g_lo, g_hi = first_index, first_index + n_asm
    for c0 in range(g_lo - g_lo % step, g_hi, step):
        gen.manual_seed(seed * 1_000_003 + c0)
        u = torch.rand(step * asm_len, device=dev, generator=gen)
        codes = torch.bucketize(u, thr).to(torch.uint8)
        # order A,C,G,T with P(C)=P(G)=gc/2: buckets [0,(1-gc)/2) A, [.., .5) C, [.5, .5+gc/2) G, rest T
        chunk = lut[codes.long()]
        del u, codes
        if n_frac > 0:
            m = torch.rand(step * asm_len, device=dev, generator=gen) < n_frac
            chunk[m] = ord("N")
            del m
        lo, hi = max(c0, g_lo), min(c0 + step, g_hi)
        out[(lo - g_lo) * asm_len : (hi - g_lo) * asm_len] = chunk[(lo - c0) * asm_len : (hi - c0) * asm_len]
        del chunk
    n_real_loci = int((~db.extra).sum() and (db.gene_locus[~db.extra].max() + 1))
    if locus_ranges is None:
        locus_ranges = ((0, n_real_loci),)
    loci = np.zeros((n_asm, len(locus_ranges)), dtype=np.int32)
    ctg_off, ctg_len, acs = [], [], [0]
    for a in range(n_asm):
        rng = np.random.default_rng(seed + first_index + a)
        taken: list[tuple[int, int]] = []
        for ci, (l0, l1) in enumerate(locus_ranges):
            li = int(rng.integers(l0, l1))
            loci[a, ci] = li
            ls = np.frombuffer(db.loci[li], dtype=np.uint8)
            ls = synth.mutate(rng, ls, float(rng.uniform(*sub)), float(rng.uniform(*indel)))
            if rng.random() < 0.5:
                ls = synth.revcomp(ls)
            while True:  # embedded loci never overlap
                pos = int(rng.integers(0, asm_len - len(ls)))
                if all(pos + len(ls) <= s0 or pos >= s1 for s0, s1 in taken):
                    break
            taken.append((pos, pos + len(ls)))
            out[a * asm_len + pos : a * asm_len + pos + len(ls)] = torch.from_numpy(ls.copy()).to(dev)
        if repeats > 0 and repeat_seq:
            rr = np.random.default_rng((seed + first_index + a) * 7919 + 13)  # its own stream: the rest of the assembly does not change
            el = np.frombuffer(repeat_seq, dtype=np.uint8)
            for _ in range(repeats):
                cp = synth.revcomp(el) if rr.random() < 0.5 else el
                while True:
                    pos = int(rr.integers(0, asm_len - len(cp)))
                    if all(pos + len(cp) <= s0 or pos >= s1 for s0, s1 in taken):
                        break
                taken.append((pos, pos + len(cp)))
                out[a * asm_len + pos : a * asm_len + pos + len(cp)] = torch.from_numpy(cp.copy()).to(dev)
        n_ctg = max(1, int(rng.poisson(mean_contigs)))
        bps = np.unique(rng.integers(1, asm_len, size=n_ctg - 1)) if n_ctg > 1 else np.zeros(0, dtype=np.int64)
        bounds = np.concatenate([[0], bps, [asm_len]]).astype(np.int64)
        ctg_off.append(a * asm_len + bounds[:-1])
        ctg_len.append(np.diff(bounds))
        acs.append(acs[-1] + len(bounds) - 1)
    return DeviceWorkload(
        ascii=out,
        contig_off=np.concatenate(ctg_off).astype(np.int64),
        contig_len=np.concatenate(ctg_len).astype(np.int32),
        asm_contig_start=np.array(acs, dtype=np.int32),
        locus=loci,
        n_asm=n_asm,
        asm_len=asm_len,
    )
