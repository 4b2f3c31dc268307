"""GPU parity for the grouped modexp / biprimality-test batch against the values recorded from the
reference's __biprime_test_v_calculation and __biprime_test_with_v_i (tests/golden)."""
from __future__ import annotations

import random

import pytest

pytestmark = pytest.mark.gpu


def _h(x: str) -> int:
    return int(x, 16)


def test_grouped_modexp_random():
    import protocols.distributed_keygen_b200 as eng

    rng = random.Random(41)
    for bits, groups, per_group in [(67, 5, 3), (130, 9, 40), (515, 4, 40), (2050, 3, 40), (2048, 2, 7), (4100, 1, 33)]:
        moduli = [rng.getrandbits(bits) | 1 | (1 << (bits - 1)) for _ in range(groups)]
        exps = [rng.getrandbits(rng.choice([bits - 3, bits // 2, 5])) for _ in range(groups)]
        exps[0] = 0
        bases = [[rng.randrange(m) for _ in range(per_group if g % 2 == 0 else max(1, per_group - 2))]
                 for g, m in enumerate(moduli)]
        bases[0][0] = 0
        got = eng.modexp_grouped(moduli, exps, bases)
        want = [[pow(b, e, m) for b in bs] for m, e, bs in zip(moduli, exps, bases)]
        assert got == want, (bits, groups)


def test_biprime_v_values_and_verdict_match_reference(biprime_vectors):
    from protocols.distributed_keygen_b200 import distributed_keygen as dkg

    cases = biprime_vectors["cases"]
    for party in (1, 2, 3):
        batch, expect = [], []
        for case in cases:
            if party > case["parties"]:
                continue
            n = _h(case["n"])
            g_values = [_h(g) for g in case["g_values"]]
            batch.append((g_values, n, _h(case["p_shares"][party - 1]), _h(case["q_shares"][party - 1])))
            expect.append([_h(v) for v in case["v"][str(party)]])
        # one grouped call per party per key width class is what compute_modulus would issue;
        # here all candidates of the same correctness parameter go together
        for correct in sorted({c["correct_param_biprime"] for c in cases}):
            idx = [i for i, c in enumerate([c for c in cases if party <= c["parties"]]) if c["correct_param_biprime"] == correct]
            got = dkg.biprime_test_v_calculation_batch([batch[i] for i in idx], party, correct)
            assert got == [expect[i] for i in idx], (party, correct)
    # verdicts from the v values of all parties
    for case in cases:
        n = _h(case["n"])
        v_by_party = {int(p): [_h(v) for v in vs] for p, vs in case["v"].items()}
        if all(len(v) >= case["correct_param_biprime"] for v in v_by_party.values()):
            assert dkg.biprime_test_with_v_i(v_by_party, n, case["correct_param_biprime"]) == case["verdict"]
    # single-candidate form, reference argument order
    case = cases[0]
    g_values = [_h(g) for g in case["g_values"]]
    v = dkg.biprime_test_v_calculation(g_values, 1, _h(case["n"]), _h(case["p_shares"][0]), _h(case["q_shares"][0]),
                                       case["correct_param_biprime"])
    assert v == [_h(x) for x in case["v"]["1"]]


def test_jacobi_and_sieve_on_gpu(biprime_vectors):
    import sympy

    import protocols.distributed_keygen_b200 as eng
    from oracle import paillier_oracle as po

    rng = random.Random(8)
    moduli, gvals = [], []
    for bits in [8, 33, 67, 130, 515, 2050, 4100]:
        for _ in range(3):
            n = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
            moduli.append(n)
            gvals.append([0, 1, n - 1, 2] + [rng.randrange(n) for _ in range(20)])
    moduli.append(3 * 5 * 7 * 11)
    gvals.append(list(range(0, 24)))
    got = eng.jacobi_batch(moduli, gvals)
    for n, gs, row in zip(moduli, gvals, got):
        assert row == [po.jacobi(g, n) for g in gs], n.bit_length()
    # the recorded reference candidates: symbols drive the same selection as the reference made
    case = biprime_vectors["cases"][3]
    n = _h(case["n"])
    gs = [_h(g) for g in case["g_values"]]
    sym = eng.jacobi_batch([n], [gs])[0]
    assert sym == [int(sympy.jacobi_symbol(g, n)) for g in gs]
    # sieve (distributed_keygen.py:1197-1209) with the reference's default prime list bound 2000
    primes = list(sympy.primerange(3, 2000))
    cands = [rng.getrandbits(2050) | 1 for _ in range(300)] + [1000003 * 999983, 1999 * (rng.getrandbits(2000) | 1), 3]
    assert eng.small_prime_sieve(cands, primes) == [po.small_prime_divisors_test(primes, c) for c in cands]


def test_biprime_round_fused_many_candidates():
    """A compute_modulus-sized round: 60 candidates x 160 g's, party 1 and party 2, against the
    oracle (Jacobi filter + selection + modexp all on the device)."""
    from oracle import keys as okeys
    from oracle import paillier_oracle as po
    from protocols.distributed_keygen_b200 import distributed_keygen as dkg

    rng = random.Random(21)
    pl, correct = 256, 40
    cands = []
    for _ in range(60):
        p_sh = [okeys.prime_candidate_share(i + 1, pl, rng) for i in range(3)]
        q_sh = [okeys.prime_candidate_share(i + 1, pl, rng) for i in range(3)]
        n = sum(p_sh) * sum(q_sh)
        gs = [rng.randint(0, n) % n for _ in range(correct * dkg.JACOBI_CORRECTION_FACTOR)]
        cands.append((gs, n, p_sh, q_sh))
    # one candidate with almost no usable g (all g = 0 -> symbol 0)
    gs0, n0, p0, q0 = cands[7]
    cands[7] = ([0] * 155 + gs0[:5], n0, p0, q0)
    for party in (1, 2):
        batch = [(gs, n, p_sh[party - 1], q_sh[party - 1]) for (gs, n, p_sh, q_sh) in cands]
        got = dkg.biprime_test_v_calculation_batch(batch, party, correct)
        want = [po.biprime_v_calculation(gs, party, n, p_i, q_i, correct) for (gs, n, p_i, q_i) in batch]
        assert got == want


def test_biprime_verdict_on_gpu(biprime_vectors):
    """Verdicts for the recorded reference candidates (real biprimes pass, the others fail), and a
    full in-process round: v values of all parties from the GPU, verdict on the GPU."""
    from oracle import keys as okeys
    from oracle import paillier_oracle as po
    from protocols.distributed_keygen_b200 import distributed_keygen as dkg

    for correct in sorted({c["correct_param_biprime"] for c in biprime_vectors["cases"]}):
        for parties in (3, 4, 5):
            cases = [c for c in biprime_vectors["cases"] if c["correct_param_biprime"] == correct and c["parties"] == parties]
            cases = [c for c in cases if all(len(v) >= correct for v in c["v"].values())]
            if not cases:
                continue
            moduli = [_h(c["n"]) for c in cases]
            v_by_party = {p: [[_h(x) for x in c["v"][str(p)]] for c in cases] for p in range(1, parties + 1)}
            assert dkg.biprime_test_with_v_i_batch(v_by_party, moduli, correct) == [c["verdict"] for c in cases]
    rng = random.Random(77)
    pl, correct = 64, 20
    cands = []
    want_bp = [True, False, True, False, False, True]
    for bp in want_bp:
        while True:
            p_sh = [okeys.prime_candidate_share(i + 1, pl, rng) for i in range(3)]
            q_sh = [okeys.prime_candidate_share(i + 1, pl, rng) for i in range(3)]
            is_bp = okeys._is_probable_prime(sum(p_sh), rng) and okeys._is_probable_prime(sum(q_sh), rng)
            if is_bp == bp:
                break
        n = sum(p_sh) * sum(q_sh)
        cands.append(([rng.randint(0, n) % n for _ in range(correct * 4)], n, p_sh, q_sh))
    v_by_party = {}
    for party in (1, 2, 3):
        batch = [(gs, n, p_sh[party - 1], q_sh[party - 1]) for (gs, n, p_sh, q_sh) in cands]
        v_by_party[party] = dkg.biprime_test_v_calculation_batch(batch, party, correct)
    got = dkg.biprime_test_with_v_i_batch(v_by_party, [c[1] for c in cands], correct)
    want = [po.biprime_verdict({p: v_by_party[p][g] for p in v_by_party}, cands[g][1], correct) for g in range(len(cands))]
    assert got == want == want_bp
    # a candidate with too few usable g's fails
    short = {p: [v_by_party[p][0][:5]] for p in v_by_party}
    assert dkg.biprime_test_with_v_i_batch(short, [cands[0][1]], correct) == [False]


def test_ragged_g_lists_stay_on_the_gpu(biprime_vectors):
    """Candidates with different numbers of g values (and one with none) in one call: padded on the
    way in, Jacobi filter and selection on the device; equals the oracle's restatement of
    distributed_keygen.py:1080-1099 per candidate."""
    from oracle import paillier_oracle as po
    from protocols.distributed_keygen_b200 import distributed_keygen as dkg

    cases = [c for c in biprime_vectors["cases"] if c["key_length"] <= 512]
    batch, want = [], []
    for k, case in enumerate(cases):
        n = _h(case["n"])
        g_values = [_h(g) for g in case["g_values"]][: max(0, len(case["g_values"]) - 13 * k)]
        p_i, q_i = _h(case["p_shares"][1]), _h(case["q_shares"][1])
        batch.append((g_values, n, p_i, q_i))
        want.append(po.biprime_v_calculation(g_values, 2, n, p_i, q_i, 20))
    batch.append(([], _h(cases[0]["n"]), _h(cases[0]["p_shares"][1]), _h(cases[0]["q_shares"][1])))
    want.append([])
    assert dkg.biprime_test_v_calculation_batch(batch, 2, 20) == want
