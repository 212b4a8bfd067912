"""Synthetic GSM C0 (BCCH carrier) capture generator - the stand-in for rtl_tcp's uint8 I,Q stream.

The reference has no recorded captures, so tests and bench.py use this seeded
generator: repeating 51-multiframe with FCCH in TS0 of frames 0,10,20,30,40 and SCH in frames
1,11,21,31,41 (64-bit extended training sequence of gsm_SCH_training_sequence_gen.m:17-19 at bit 42),
random bursts elsewhere, GMSK BT 0.3 / L=4 / h=0.5 at 8 samples per symbol and 1250 samples per slot,
sampling-clock error applied by exact re-timing of the continuous phase, carrier error, AWGN, DC, and
round+clip to uint8 around 127.5 - the wire format `fread(tcp, n, 'uint8')` sees (gsm_sync_demod.m:96).

Written with torch so the same code runs on the CPU (tests, fixtures) and on the GPU (bench.py, where
1024 x 21.7 M samples would take far too long on the host).  torch here is data plumbing only.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

SYMBOL_RATE = (1625.0 / 6.0) * 1e3
SLOT = 1250                 # samples per slot at 8x
SYM_PER_SLOT = 156          # impulses per slot (148 burst bits + 8 guard bits; the quarter bit is 2 samples of hold)
FRAME = 8 * SLOT
MULTIFRAME = 51 * FRAME

SCH_TRAINING_BITS = (1, 0, 1, 1, 1, 0, 0, 1, 0, 1, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0,
                     0, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 1, 0, 1, 1, 0, 1, 0, 1, 0, 0, 0,
                     1, 0, 1, 0, 1, 1, 1, 0, 1, 1, 0, 0, 0, 0, 1, 1, 0, 1, 1)

_SIGMA = math.sqrt(math.log(2.0)) / (2.0 * math.pi * 0.3)


def _big_f(u: torch.Tensor) -> torch.Tensor:
    return u * 0.5 * torch.erfc(-u / math.sqrt(2.0)) + torch.exp(-0.5 * u * u) / math.sqrt(2.0 * math.pi)


def _big_g(t: torch.Tensor) -> torch.Tensor:
    return (_SIGMA / 2.0) * (_big_f((t + 0.5) / _SIGMA) - _big_f((t - 0.5) / _SIGMA))


def gmsk_q(tau: torch.Tensor) -> torch.Tensor:
    """Integrated, L=4-truncated GMSK frequency pulse; tau in symbols; q(<=0)=0, q(>=4)=1."""
    tau = tau.clamp(0.0, 4.0)
    c = torch.tensor([-2.0, 2.0], dtype=tau.dtype, device=tau.device)
    g = _big_g(c)
    return (_big_g(tau - 2.0) - g[0]) / (g[1] - g[0])


@dataclass
class StreamSpec:
    """Impairments of one synthetic dongle stream."""
    seed: int
    n_samples: int
    sampling_ppm: float = 0.0       # receiver clock fast by this much -> burst spacing grows by (1+e)
    carrier_ppm: float = 0.0        # of carrier_freq
    carrier_freq: float = 957.4e6   # gsm_sync_demod.m:14
    snr_db: float = 20.0
    phase0: float = 0.0
    start_offset: float = 0.0       # nominal samples into the multiframe at receiver sample 0
    amplitude: float = 40.0
    dc: complex = 127.4 + 127.6j
    drop_fcch: tuple = field(default_factory=tuple)   # indices (0-based, in order of appearance) of FCCH bursts to replace by data
    noise_only: bool = False
    tsc: int | None = None          # normal training sequence code (0..7) put at bits 61..86 of every normal burst; None = random bits


def random_spec(seed: int, n_samples: int, carrier_freq: float = 957.4e6) -> StreamSpec:
    """The impairment distribution of SURVEY.md section 8(d), narrowed to the range the reference chain locks on."""
    rng = np.random.default_rng(seed)
    return StreamSpec(seed=seed, n_samples=n_samples,
                      sampling_ppm=float(rng.uniform(-40.0, 35.0)),
                      carrier_ppm=float(rng.uniform(-25.0, 25.0)),
                      carrier_freq=carrier_freq,
                      snr_db=float(rng.uniform(15.0, 25.0)),
                      phase0=float(rng.uniform(0.0, 2.0 * math.pi)),
                      start_offset=float(rng.integers(0, MULTIFRAME)))


NORMAL_TRAINING_BITS = (      # gsm_normal_training_sequence_gen.m:17-24
    (0, 0, 1, 0, 0, 1, 0, 1, 1, 1, 0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 1, 0, 1, 1, 1),
    (0, 0, 1, 0, 1, 1, 0, 1, 1, 1, 0, 1, 1, 1, 1, 0, 0, 0, 1, 0, 1, 1, 0, 1, 1, 1),
    (0, 1, 0, 0, 0, 0, 1, 1, 1, 0, 1, 1, 1, 0, 1, 0, 0, 1, 0, 0, 0, 0, 1, 1, 1, 0),
    (0, 1, 0, 0, 0, 1, 1, 1, 1, 0, 1, 1, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 1, 1, 0),
    (0, 0, 0, 1, 1, 0, 1, 0, 1, 1, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1, 1, 0, 1, 0, 1, 1),
    (0, 1, 0, 0, 1, 1, 1, 0, 1, 0, 1, 1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 1, 1, 0, 1, 0),
    (1, 0, 1, 0, 0, 1, 1, 1, 1, 1, 0, 1, 1, 0, 0, 0, 1, 0, 1, 0, 0, 1, 1, 1, 1, 1),
    (1, 1, 1, 0, 1, 1, 1, 1, 0, 0, 0, 1, 0, 0, 1, 0, 1, 1, 1, 0, 1, 1, 1, 1, 0, 0))


def _slot_bits(n_slots: int, gen: torch.Generator, device, drop_fcch=(), tsc=None) -> torch.Tensor:
    """[n_slots, 156] bits: TS0 of frames 0/10/20/30/40 (mod 51) FCCH, 1/11/21/31/41 SCH, else random."""
    bits = torch.randint(0, 2, (n_slots, SYM_PER_SLOT), generator=gen, device=device, dtype=torch.int64)
    bits[:, 0:3] = 0
    bits[:, 145:148] = 0
    bits[:, 148:156] = 1
    if tsc is not None:
        bits[:, 61:87] = torch.tensor(NORMAL_TRAINING_BITS[int(tsc)], device=device, dtype=torch.int64)
    slot = torch.arange(n_slots, device=device)
    frame = (slot // 8) % 51
    ts0 = (slot % 8) == 0
    is_fcch = ts0 & (frame % 10 == 0) & (frame <= 40)
    is_sch = ts0 & (frame % 10 == 1) & (frame <= 41)
    fcch_slots = torch.nonzero(is_fcch).flatten()
    if len(drop_fcch):
        keep = torch.ones(len(fcch_slots), dtype=torch.bool, device=device)
        for d in drop_fcch:
            if d < len(keep):
                keep[d] = False
        fcch_slots = fcch_slots[keep]
    bits[fcch_slots, 0:148] = 0
    ts = torch.tensor(SCH_TRAINING_BITS, device=device, dtype=torch.int64)
    bits[is_sch, 42:106] = ts
    return bits


def generate_stream(spec: StreamSpec, device="cpu", out: torch.Tensor | None = None) -> torch.Tensor:
    """Returns 2*n_samples uint8 (I,Q interleaved).  `out` may be a preallocated uint8 view to fill."""
    dev = torch.device(device)
    n = spec.n_samples
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(spec.seed))
    f64 = torch.float64
    inv = 1.0 / (1.0 + spec.sampling_ppm * 1e-6)
    t_end = spec.start_offset + n * inv
    n_slots = int(t_end // SLOT) + 3
    bits = _slot_bits(n_slots, gen, dev, spec.drop_fcch, spec.tsc).flatten()
    prev = torch.cat([torch.ones(1, dtype=torch.int64, device=dev), bits[:-1]])
    a = (1 - 2 * (bits ^ prev)).to(torch.int64)                 # +1 where a bit equals its predecessor
    psum = torch.cumsum(a, 0)                                   # P[m] = sum_{i<=m} a_i

    result = out if out is not None else torch.empty(2 * n, dtype=torch.uint8, device=dev)
    noise_sigma = spec.amplitude / math.sqrt(2.0) * 10.0 ** (-spec.snr_db / 20.0)
    w_off = 2.0 * math.pi * (spec.carrier_ppm * 1e-6 * spec.carrier_freq) / (SYMBOL_RATE * 8)
    chunk = 1 << 22
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        idx = torch.arange(c0, c1, device=dev, dtype=f64)
        t = spec.start_offset + idx * inv                        # nominal time in samples
        slot = torch.floor(t / SLOT)
        w = t - slot * SLOT
        k = torch.clamp(torch.floor(w / 8.0), max=SYM_PER_SLOT - 1)
        g = (slot * SYM_PER_SLOT + k).to(torch.int64)
        gm4 = g - 4
        base = torch.where(gm4 >= 0, psum[gm4.clamp(min=0)], torch.zeros_like(gm4))
        phase = (base % 4).to(f64)                               # whole quarter turns, exact
        for j in range(4):
            gj = g - j
            valid = gj >= 0
            gjc = gj.clamp(min=0)
            tj = (gjc // SYM_PER_SLOT).to(f64) * SLOT + (gjc % SYM_PER_SLOT).to(f64) * 8.0
            phase = phase + torch.where(valid, a[gjc].to(f64) * gmsk_q((t - tj) / 8.0), torch.zeros_like(t))
        theta = (math.pi / 2.0) * phase + (idx * w_off + spec.phase0)
        amp = 0.0 if spec.noise_only else spec.amplitude
        re = amp * torch.cos(theta) + noise_sigma * torch.randn(c1 - c0, generator=gen, device=dev, dtype=f64)
        im = amp * torch.sin(theta) + noise_sigma * torch.randn(c1 - c0, generator=gen, device=dev, dtype=f64)
        re = torch.clamp(torch.round(re + spec.dc.real), 0, 255).to(torch.uint8)
        im = torch.clamp(torch.round(im + spec.dc.imag), 0, 255).to(torch.uint8)
        result[2 * c0:2 * c1:2] = re
        result[2 * c0 + 1:2 * c1:2] = im
    return result


def generate_batch(specs, device="cpu") -> torch.Tensor:
    """[D, 2N] uint8, one row per stream (== one column of the reference's 2N x D matrix)."""
    n = specs[0].n_samples
    out = torch.empty((len(specs), 2 * n), dtype=torch.uint8, device=device)
    for d, sp in enumerate(specs):
        assert sp.n_samples == n
        generate_stream(sp, device, out[d])
    return out


def true_fcch_starts(spec: StreamSpec) -> np.ndarray:
    """1-based receiver sample index of every FCCH burst start inside the capture (sanity checks only)."""
    scale = 1.0 + spec.sampling_ppm * 1e-6
    res = []
    t_end = spec.start_offset + spec.n_samples / scale
    mf0 = int(spec.start_offset // MULTIFRAME)
    m = mf0
    while m * MULTIFRAME < t_end:
        for f in (0, 10, 20, 30, 40):
            t = m * MULTIFRAME + f * FRAME
            nidx = (t - spec.start_offset) * scale
            if 0 <= nidx < spec.n_samples:
                res.append(nidx + 1)
        m += 1
    return np.array(res)
