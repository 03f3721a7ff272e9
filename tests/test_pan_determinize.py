"""CPU: the Pan determinization sampler (new behaviour; the reference has none -- SURVEY.md 8a P5).  Parity is by
construction properties: every sample is a complete-information state consistent with the player's view."""
import numpy as np
import pytest
import gameai_b200 as g

M48 = (1 << 48) - 1


def fields(s):
    stack = int(s[0]) & M48
    hands = [(int(s[1 + p]) >> 4) & M48 for p in range(4)]
    counts = [int(s[1 + p]) & 15 for p in range(4)]
    return stack, hands, counts


def crisp(c):
    return (((c >> 1) ^ c) & 0x555555555555) == 0


@pytest.mark.parametrize("npl", [2, 3, 4])
def test_samples_are_consistent_with_the_view(npl):
    from oracle import pan
    po = pan.port()
    S = po.gen_reachable(npl, 60, 31 + npl, 40)
    checked = 0
    for i, full in enumerate(S):
        player = i % npl
        fstack, fhands, fcounts = fields(full)
        if sum(fcounts[:npl]) + bin(fstack).count("1") // 2 != 24:
            continue                                               # a 4-bit count wrapped (2-player hands > 15 cards)
        view = po.player_known_state(npl, full, player)
        D = g.pan_determinize(view, npl, player, 16, key=99, first_id=1000 * i)
        assert (g.pan_determinize(view, npl, player, 16, key=99, first_id=1000 * i) == D).all()        # pure function of (view, key, id)
        seen = set()
        for d in D:
            stack, hands, counts = fields(d)
            assert stack == fstack and hands[player] == fhands[player] and counts[:npl] == fcounts[:npl]
            assert int(d[0]) >> 48 == int(full[0]) >> 48                                             # current player / terminal flag kept
            allc = stack
            for p in range(npl):
                assert crisp(hands[p]) and bin(hands[p]).count("1") == 2 * counts[p]
                assert allc & hands[p] == 0
                allc |= hands[p]
            assert allc == M48                                                                        # all 24 cards accounted for once
            seen.add(tuple(hands))
            c, _ = po.movegen_batch(npl, d[None])
            assert c[0] >= 1 or po.terminal_batch(npl, d[None])[0][0] == 1
        if npl > 2 and sum(fcounts[p] for p in range(npl) if p != player and fcounts[p]) > 3 and sum(1 for p in range(npl) if p != player and fcounts[p]) > 1:
            assert len(seen) > 1                                                                      # genuinely different deals
        if npl == 2:
            assert (D == D[0]).all() and fields(D[0])[1][1 - player] == fhands[1 - player]             # two players: nothing hidden
        checked += 1
    assert checked > 30


def test_rejects_inconsistent_views():
    from oracle import pan
    po = pan.port()
    full = po.gen_reachable(3, 1, 5, 10)[0]
    view = po.player_known_state(3, full, 0)
    bad = view.copy(); bad[2] = (int(bad[2]) & ~15) | ((int(bad[2]) & 15) + 1)       # claim one more card than exists
    with pytest.raises(g.GaiError):
        g.pan_determinize(bad, 3, 0, 4)
    with pytest.raises(g.GaiError):
        g.pan_determinize(view, 3, 1, 4)                                              # player 1's own hand is fuzzy in player 0's view


def test_cards_an_opponent_was_seen_taking_stay_in_that_hand():
    """After a take the taker's cards are certain (11) in everybody's view (apply_move, GraWPanaZasadyV2.cpp:541-545, which
    UpdatePlayerKnownState :313-324 runs on the view): every sampled deal must leave them in the taker's hand."""
    from oracle import pan
    po = pan.port()
    found = 0
    for npl in (3, 4):
        S = po.gen_reachable(npl, 400, 77 + npl, 40)
        for i, full in enumerate(S):
            c, m = po.movegen_batch(npl, full[None])
            takes = [int(m[0, k]) for k in range(int(c[0])) if ((int(m[0, k]) >> 51) & 7) == 1]      # Move::take_cards
            if not takes:
                continue
            mover = (int(full[0]) >> 48) & 7
            observer = (mover + 1) % npl
            fstack, fhands, fcounts = fields(full)
            if sum(fcounts[:npl]) + bin(fstack).count("1") // 2 != 24:
                continue
            view = po.player_known_state(npl, full, observer)
            mv = np.array([takes[0]], dtype=np.uint64)
            view2 = po.apply_batch(npl, view[None], mv)[0]           # what UpdatePlayerKnownState does to the view
            full2 = po.apply_batch(npl, full[None], mv)[0]
            taken = takes[0] & M48
            assert crisp(taken) and taken
            s2, _, c2 = fields(full2)
            if sum(c2[:npl]) + bin(s2).count("1") // 2 != 24:
                continue                                               # the taker's 4-bit count wrapped past 15
            D = g.pan_determinize(view2, npl, observer, 32, key=5, first_id=i)
            for d in D:
                stack, hands, counts = fields(d)
                assert hands[mover] & taken == taken                  # the seen cards are where they were seen going
                assert stack == fields(full2)[0] and counts[:npl] == fields(full2)[2][:npl]
                allc = stack
                for p in range(npl):
                    assert crisp(hands[p]) and bin(hands[p]).count("1") == 2 * counts[p] and allc & hands[p] == 0
                    allc |= hands[p]
                assert allc == M48
            found += 1
    assert found > 20
