"""The shared-memory layouts of the fused engines' exchanges (csrc/sds_packed.cuh), restated in Python:
every store and load pattern of the 8-elements-per-thread Stockham schedule must be free of bank
conflicts AND additive over the schedule's index split (so that an access is [register + immediate]).
CPU-only: these are the 'checked exhaustively' claims of the kernel comments."""
import pytest


def npass(n):
    return (n.bit_length() - 1 + 2) // 3


def radix(n, p):
    lg = n.bit_length() - 1 - 3 * p
    return 8 if lg >= 3 else 1 << lg


def ns(n, p):
    out = 1
    for i in range(p):
        out *= radix(n, i)
    return out


def il_pad(j):  # 16-byte (data, noise) slots: TeamFftX::il_pad
    return j + (j >> 3)


def p1_pad(j):  # 8-byte slots of the single transform: TeamFftX::p1_pad
    return j + (j >> 4)


def patterns(n):
    """(kind, lanes -> slot index) for every exchange access of a length-n transform."""
    tpr = n // 8
    for p in range(npass(n) - 1):
        NS = ns(n, p)
        for q in range(8):
            yield "store", [(tau // NS) * (NS * 8) + tau % NS + q * NS for tau in range(tpr)]
    for r in range(8):
        yield "load", [tau + tpr * r for tau in range(tpr)]


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048])
def test_interleaved_layout_is_conflict_free_per_quarter_warp(n):
    # a 128-bit access is served per quarter warp: its 8 slots must sit on 8 distinct 16-byte bank groups
    for kind, idx in patterns(n):
        lanes = idx if len(idx) >= 32 else idx * (32 // len(idx))  # several rows per warp: same pattern per row
        for w0 in range(0, len(lanes), 8):
            groups = {il_pad(j) % 8 for j in lanes[w0:w0 + 8]}
            assert len(groups) == 8, (n, kind, w0)


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048])
def test_interleaved_layout_is_additive(n):
    tpr = n // 8
    for p in range(npass(n) - 1):
        NS = ns(n, p)
        for tau in range(tpr):
            base = (tau // NS) * (NS * 8) + tau % NS
            for q in range(8):
                assert il_pad(base + q * NS) == il_pad(base) + il_pad(q * NS)
    for tau in range(tpr):
        for r in range(8):
            assert il_pad(tau + tpr * r) == il_pad(tau) + il_pad(tpr * r)


def test_single_transform_layout_n256():
    # 64-bit accesses are served per half warp: 16 distinct 8-byte slots modulo 16; only pass 0 of the
    # 8 x 8 x 4 schedule goes through shared memory (the last exchange is in tensor memory)
    n, tpr = 256, 32
    for h in range(2):
        taus = range(16 * h, 16 * h + 16)
        for q in range(8):
            assert len({p1_pad(8 * tau + q) % 16 for tau in taus}) == 16
        for r in range(8):
            assert len({p1_pad(tau + tpr * r) % 16 for tau in taus}) == 16
    for tau in range(tpr):
        for q in range(8):
            assert p1_pad(8 * tau + q) == p1_pad(8 * tau) + p1_pad(q)
            assert p1_pad(tau + tpr * q) == p1_pad(tau) + p1_pad(tpr * q)


def tm_tau(lane):  # the team member a lane acts as after the tensor-memory exchange
    return ((lane & 3) << 3) | (lane >> 2)


def test_tensor_memory_exchange_is_the_stockham_exchange_renamed():
    """Last exchange of the 8 x 8 x 4 schedule at N = 256: output t of thread tau goes to position
    64 (tau >> 3) + (tau & 7) + 8 t, which thread pos & 31 reads as slot pos >> 5.  The hardware pair
    tcgen05.st.32x32b / tcgen05.ld.16x256b (profiles/r2_tmem_exchange_microbench.txt) delivers
    (source thread tau, value t) to thread 4 (tau & 7) + (t & 3), loaded slot (tau >> 3 & 1) + 2 (t >> 2) + 4 (tau >> 4).
    With the lane renaming tm_tau and the loaded slots 1 <-> 2, 5 <-> 6 swapped the two agree."""
    swap = {0: 0, 1: 2, 2: 1, 3: 3, 4: 4, 5: 6, 6: 5, 7: 7}
    for tau in range(32):
        for t in range(8):
            pos = 64 * (tau >> 3) + (tau & 7) + 8 * t
            want_thread, want_slot = pos & 31, pos >> 5
            lane = 4 * (tau & 7) + (t & 3)
            loaded = ((tau >> 3) & 1) + 2 * (t >> 2) + 4 * (tau >> 4)
            assert tm_tau(lane) == want_thread
            assert swap[loaded] == want_slot
    assert sorted(tm_tau(l) for l in range(32)) == list(range(32))
