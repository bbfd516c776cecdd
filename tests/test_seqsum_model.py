"""CPU model of seq_sum_scan_kernel (spfresh_b200/csrc/ops.cu): the strictly sequential f32 fold of
hierarchical.rs:278 computed per binade with integer arithmetic and a two-state (parity) transducer.
The model restates the kernel's arithmetic in plain Python so the algorithm is pinned against
numpy's sequential float32 cumsum without a GPU; the kernel itself is checked against the serial
add chain in tests/test_gpu_parity.py::test_sequential_sum_scan_is_bit_exact."""
import numpy as np
import pytest

LIMIT = 1 << 24


def decode(xb, eeff):
    """floor(x/u), rounds-up flag, tie flag for ulp u = 2^(eeff - 150)."""
    ex = (xb >> 23) & 0xFF
    mx = ((xb & 0x7FFFFF) | 0x800000) if ex else (xb & 0x7FFFFF)
    exe = ex if ex else 1
    if exe > eeff:
        return LIMIT, 0, 0
    if exe == eeff:
        return mx, 0, 0
    sh = eeff - exe
    if sh > 25:
        return 0, 0, 0
    rem, half = mx & ((1 << sh) - 1), 1 << (sh - 1)
    return mx >> sh, int(rem > half), int(rem == half)


def to_bits(a, eeff):
    return a if a < 0x800000 else ((eeff << 23) | (a & 0x7FFFFF))


def model_sum(v, run=5):
    """Windows of `run`-element runs composed as transducers, restart at every binade crossing."""
    bits = v.view(np.uint32).tolist()
    n, pos, acc = len(bits), 0, 0
    window = 7 * run
    while pos < n and (acc >> 23) != 255:
        ea = acc >> 23
        eeff = ea if ea else 1
        a0 = ((acc & 0x7FFFFF) | 0x800000) if ea else acc
        dec = [decode(bits[i], eeff) if i < n else (0, 0, 0) for i in range(pos, pos + window)]
        # run summaries for both parities, composed left to right (the scan of the kernel)
        summaries = []
        for r0 in range(0, window, run):
            out = []
            for p in (0, 1):
                k_sum, par = 0, p
                for f, gt, tie in dec[r0:r0 + run]:
                    kk = f + (((par + f) & 1) if tie else gt)
                    k_sum += kk
                    par = (par + kk) & 1
                out.append((k_sum, par))
            summaries.append(out)
        par, a_in, starts = a0 & 1, a0, []
        for sm in summaries:
            starts.append(a_in)
            k_sum, par = sm[par]
            a_in += k_sum
        # second walk: first element that reaches the next binade
        cross = None
        for ri, a in enumerate(starts):
            if a >= LIMIT:
                break
            for j, (f, gt, tie) in enumerate(dec[ri * run:(ri + 1) * run]):
                kk = f + (((a + f) & 1) if tie else gt)
                if a + kk >= LIMIT:
                    cross = (ri * run + j, a)
                    break
                a += kk
            if cross:
                break
        if cross is None:
            acc = to_bits(a_in, eeff)
            pos += window
        else:
            ci, a = cross
            before = np.array([to_bits(a, eeff)], np.uint32).view(np.float32)[0]
            with np.errstate(over="ignore"):
                acc = int(np.array([before + v[pos + ci]], np.float32).view(np.uint32)[0])
            pos += ci + 1
    out = np.array([acc], np.uint32).view(np.float32)[0]
    for i in range(pos, n):                      # tail after an overflow to +inf
        out = np.float32(out + v[i])
    return out


@pytest.mark.parametrize("case", range(10))
def test_transducer_model_equals_sequential_cumsum(case):
    rng = np.random.default_rng(100 + case)
    n = int(rng.integers(1, 700))
    v = [rng.random(n, dtype=np.float32) * 300.0,
         (rng.integers(0, 4096, n) * 0.25).astype(np.float32),                       # exact ties
         np.exp(rng.normal(0, 12, n)).astype(np.float32),                            # 30 binades
         np.concatenate([np.zeros(20, np.float32), rng.random(n, dtype=np.float32) * 1e-41]),   # denormals
         np.concatenate([[1.0], np.full(n, 2.0 ** -24, np.float32)]).astype(np.float32),        # half-ulp ties
         np.concatenate([[1.0], np.full(n, 2.0 ** -24 * 1.5, np.float32)]).astype(np.float32),
         np.full(n, 16777216.0, np.float32),
         np.full(min(n, 40), 3.0e38, np.float32),                                    # overflow
         (rng.integers(0, 3, n)).astype(np.float32),
         np.concatenate([rng.random(n, dtype=np.float32) * 1e-38, rng.random(n, dtype=np.float32)])][case]
    v = np.ascontiguousarray(v, np.float32)
    with np.errstate(over="ignore"):
        want = np.cumsum(v, dtype=np.float32)[-1]
    got = model_sum(v)
    assert np.array([got]).view(np.uint32)[0] == np.array([want]).view(np.uint32)[0], (case, n, got, want)


SAT = 1 << 26


def compose(a, b):
    """Transducer a followed by b.  A transducer is (k0, k1): the sum of the rounded quotients when the
    run is entered with parity 0 / 1, saturated at 2^26; the parity behind the run is (p + k[p]) & 1."""
    return (min(a[0] + b[a[0] & 1], SAT), min(a[1] + b[(1 + a[1]) & 1], SAT))


IDENT = (0, 0)


def model_sum_cluster(v, run=4, lanes=3, warps=2, ctas=3, prefix=5):
    """seq_sum_cluster_kernel: the first `prefix` elements by the plain add chain; then fixed windows of
    ctas x warps x lanes runs; a crossing restarts inside the SAME window with the elements in front of
    the restart position masked to +0; summaries are composed thread -> warp -> CTA -> cluster and only
    the first crossing CTA walks its runs again."""
    bits = v.view(np.uint32).tolist()
    n = len(bits)
    pre = np.float32(0)
    with np.errstate(over="ignore"):
        for i in range(min(n, prefix)):
            pre = np.float32(pre + v[i])
    acc, start, wbase = int(np.array([pre], np.float32).view(np.uint32)[0]), min(n, prefix), 0
    per_cta = warps * lanes * run
    window = ctas * per_cta
    while start < n and (acc >> 23) != 255:
        ea = acc >> 23
        eeff = ea if ea else 1
        a0 = ((acc & 0x7FFFFF) | 0x800000) if ea else acc
        dec = [decode(bits[i] if start <= i < n else 0, eeff) for i in range(wbase, wbase + window)]
        thread = []
        for r0 in range(0, window, run):
            out = []
            for p in (0, 1):
                k_sum, par = 0, p
                for f, gt, tie in dec[r0:r0 + run]:
                    kk = f + (((par + f) & 1) if tie else gt)
                    k_sum += kk
                    par = (par + kk) & 1
                out.append(min(k_sum, SAT))
            thread.append(tuple(out))
        # exclusive prefixes per level
        cta_tot, warp_ex, lane_ex = [], [], []
        for c in range(ctas):
            wacc = IDENT
            for w in range(warps):
                warp_ex.append(wacc)
                lacc = IDENT
                for l in range(lanes):
                    lane_ex.append(lacc)
                    lacc = compose(lacc, thread[(c * warps + w) * lanes + l])
                wacc = compose(wacc, lacc)
            cta_tot.append(wacc)
        p0 = a0 & 1
        cinc, inc = [], IDENT
        for c in range(ctas):
            inc = compose(inc, cta_tot[c])
            cinc.append(inc[p0])
        rc = next((c for c in range(ctas) if a0 + cinc[c] >= LIMIT), None)
        if rc is None:
            acc = to_bits(a0 + cinc[-1], eeff)
            wbase += window
            start = wbase
            continue
        best = None
        for w in range(warps):
            for l in range(lanes):
                t = (rc * warps + w) * lanes + l
                a = a0 + (cinc[rc - 1] if rc else 0)
                a = min(a + warp_ex[rc * warps + w][a & 1], SAT)
                a = min(a + lane_ex[t][a & 1], SAT)
                mine = thread[t][a & 1]
                if a < LIMIT and a + mine >= LIMIT:
                    for j, (f, gt, tie) in enumerate(dec[t * run:(t + 1) * run]):
                        kk = f + (((a + f) & 1) if tie else gt)
                        if a + kk >= LIMIT:
                            key = (t * run + j, a)
                            best = key if best is None or key < best else best
                            break
                        a += kk
        ci, a = best
        assert wbase + ci >= start
        before = np.array([to_bits(a, eeff)], np.uint32).view(np.float32)[0]
        with np.errstate(over="ignore"):
            acc = int(np.array([before + v[wbase + ci]], np.float32).view(np.uint32)[0])
        start = wbase + ci + 1
        if start >= wbase + window:
            wbase += window
    out = np.array([acc], np.uint32).view(np.float32)[0]
    for i in range(start, n):
        out = np.float32(out + v[i])
    return out


@pytest.mark.parametrize("case", range(10))
def test_cluster_model_equals_sequential_cumsum(case):
    rng = np.random.default_rng(300 + case)
    n = int(rng.integers(1, 900))
    v = [rng.random(n, dtype=np.float32) * 300.0,
         (rng.integers(0, 4096, n) * 0.25).astype(np.float32),
         np.exp(rng.normal(0, 12, n)).astype(np.float32),
         np.concatenate([np.zeros(20, np.float32), rng.random(n, dtype=np.float32) * 1e-41]),
         np.concatenate([[1.0], np.full(n, 2.0 ** -24, np.float32)]).astype(np.float32),
         np.concatenate([[1.0], np.full(n, 2.0 ** -24 * 1.5, np.float32)]).astype(np.float32),
         np.full(n, 16777216.0, np.float32),
         np.full(min(n, 40), 3.0e38, np.float32),
         (rng.integers(0, 3, n)).astype(np.float32),
         np.concatenate([rng.random(n, dtype=np.float32) * 1e-38, rng.random(n, dtype=np.float32)])][case]
    v = np.ascontiguousarray(v, np.float32)
    with np.errstate(over="ignore"):
        want = np.cumsum(v, dtype=np.float32)[-1]
    got = model_sum_cluster(v)
    assert np.array([got]).view(np.uint32)[0] == np.array([want]).view(np.uint32)[0], (case, n, got, want)
