"""Model check (CPU, exhaustive over interleavings) of the tile-end hand-over of the fp16 attention kernel
(openglue_b200/csrc/attention_f16t.cuh, kernel template parameter SWAP / launcher attention_f16t_launch_t).

With SWAP the team that owns a tile's last key block merges the tile behind a waiting ``bar.sync 3, 512`` and the other team deposits its
partial result behind a non-waiting ``bar.arrive 3, 512``.  One named barrier serves every tile, so the pairing is sound only if the
team that deposits for tile t+1 can never reach the barrier before the team that merges tile t has reached it - otherwise one barrier
generation would be completed by two arrivals of the SAME team and the merger would later pair up with a deposit of the wrong tile.

The model keeps exactly the orderings the kernel enforces and nothing else (no timing):
  * a team processes its key blocks (global block index parity = team) in order, then the tile end, then the next tile;
  * a block can be processed once its QK^T has been issued; the issuer works in global block order, QK^T(g) waits for P(g-2) (the
    same team's buffer) and, for the first block of tile t >= 1, for the tile's Q, which the owner of tile t-1's last block writes when
    the logits of that last block are there (before it processes the block);
  * ``bar.sync`` blocks until its barrier generation is complete, ``bar.arrive`` does not; a generation is complete after two
    team-arrivals (2 x 256 threads), whoever they come from.
Every reachable interleaving is explored.  Result (asserted below): the pairing is order-safe from five key blocks per tile on (a
team's third block of a tile needs a QK^T that is issued behind one waiting for the other team's first P of that tile, which that team
writes after the previous tile's barrier) and NOT order-safe for 1 - 4 blocks, where a team can run through a whole tile while the
other one has not reached the previous tile's barrier yet - which is why the launcher uses SWAP only from six key blocks per tile on
and the symmetric form (both teams wait) below that.
"""
import os
import re

import pytest


def explore(nblk: int, tiles: int, swap: bool):
    """-> (violation reachable?, number of states).  State = (pc of team 0, pc of team 1, QK^T issued, open barrier generation)."""
    total = nblk * tiles
    progs = ([], [])
    blk_step, q_step = {}, {}                       # global block -> (team, index of its step); tile -> (team, index of its Q write)
    for t in range(tiles):
        it0 = t * nblk
        last = it0 + nblk - 1
        for g in range(it0, it0 + nblk):
            team = g & 1
            if g == last and t + 1 < tiles:         # writer_next: the owner of the tile's last block writes the next tile's Q
                q_step[t + 1] = (team, len(progs[team]))
                progs[team].append(('q', t + 1, g))
            blk_step[g] = (team, len(progs[team]))
            progs[team].append(('blk', g))
        merger = (last & 1) if swap else 0
        for team in (0, 1):
            progs[team].append(('sync' if (not swap or team == merger) else 'arrive', t))

    def past(pc, where):                            # has the team completed the step `where` = (team, index)?
        return pc[where[0]] > where[1]

    def adv(pc, team):
        return (pc[0] + 1, pc[1]) if team == 0 else (pc[0], pc[1] + 1)

    bad = False
    seen = set()
    stack = [((0, 0), 0, ())]
    while stack:
        state = stack.pop()
        if state in seen:
            continue
        seen.add(state)
        pc, issued, bar = state
        # --- the QK^T issuer (one thread, global block order)
        if issued < total:
            g = issued
            t = g // nblk
            ok = g < 2 or past(pc, blk_step[g - 2])
            if ok and t > 0 and g == t * nblk:
                ok = past(pc, q_step[t])
            if ok:
                stack.append((pc, issued + 1, bar))
        # --- the two softmax teams
        for team in (0, 1):
            i = pc[team]
            if i >= len(progs[team]) or any(a[0] == team and a[2] == 'sync' for a in bar):
                continue                            # finished, or blocked in bar.sync
            st = progs[team][i]
            if st[0] == 'q':                        # needs the logits of the tile's last block (and with them q_free)
                if st[2] < issued:
                    stack.append((adv(pc, team), issued, bar))
            elif st[0] == 'blk':
                if st[1] < issued:
                    stack.append((adv(pc, team), issued, bar))
            else:
                nbar = bar + ((team, st[1], st[0]),)
                if len(nbar) < 2:
                    # bar.arrive moves on at once; bar.sync registers and blocks (pc stays, the entry in `bar` marks it as waiting)
                    stack.append((adv(pc, team) if st[0] == 'arrive' else pc, issued, nbar))
                else:                               # the generation is complete
                    (ta, tile_a, _), (tb, tile_b, _) = nbar
                    if ta == tb or tile_a != tile_b:
                        bad = True                  # two arrivals of one team, or deposits / merges of different tiles paired up
                    npc = adv(pc, team)
                    other = nbar[0]
                    if other[2] == 'sync' and other[0] != team:
                        npc = adv(npc, other[0])    # the waiting team is released
                    stack.append((npc, issued, ()))
    return bad, len(seen)


@pytest.mark.parametrize('nblk', [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 32])
def test_symmetric_form_is_always_safe(nblk):
    bad, n = explore(nblk, 4 if nblk < 32 else 3, swap=False)
    assert not bad and n > 0


@pytest.mark.parametrize('nblk', [5, 6, 7, 8, 9, 10, 12, 32])
def test_role_swap_is_order_safe_where_the_launcher_uses_it(nblk):
    bad, n = explore(nblk, 4 if nblk < 32 else 3, swap=True)
    assert not bad and n > 0


@pytest.mark.parametrize('nblk', [1, 2, 3, 4])
def test_role_swap_rests_on_timing_for_short_tiles(nblk):
    """The reason for the launcher's threshold: with up to four key blocks per tile an interleaving exists (in the order-only model) in
    which one team reaches the barrier twice (tile t, then tile t+1) before the other team has reached it for tile t."""
    bad, _ = explore(nblk, 4, swap=True)
    assert bad


def test_launcher_threshold_matches_the_model():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, 'openglue_b200', 'csrc', 'attention_f16t.cuh')).read()
    m = re.search(r'OG_ATTN_MERGER_LAST && nblk >= (\d+)', src)
    assert m and int(m.group(1)) >= 5
