"""Weighted row-tile partition of the packed multi-GPU gather (make_part_table / make_params in csrc/rtb200.cu), mirrored here in Python:
the table must give every slot exactly one owner, every rank != 0 the same share, and spread rank 0's slots over the period. The GPU
tests check the library's own table through whole frames rendered into a poisoned framebuffer (a tile nobody renders shows up)."""
import pytest


def make_part_table(world, sink, peer):
    if world < 2:
        return [0]
    if sink < 0 or peer < 1 or sink > 16 or peer > 16 or sink + peer * (world - 1) > 64 or world > 8:
        sink, peer = 1, 1
    period = sink + peer * (world - 1)
    owner, nxt = [], 1
    for j in range(period):
        is_sink = ((j + 1) * sink) // period > (j * sink) // period
        if is_sink:
            owner.append(0)
        else:
            owner.append(nxt); nxt = nxt + 1 if nxt + 1 < world else 1
    return owner


@pytest.mark.parametrize("world,sink,peer", [(2, 1, 1), (4, 4, 5), (8, 1, 2), (8, 0, 1), (3, 0, 4), (8, 3, 5), (4, 2, 3), (8, 16, 6), (5, 1, 1)])
def test_every_slot_has_one_owner_and_peers_get_equal_shares(world, sink, peer):
    owner = make_part_table(world, sink, peer)
    if sink + peer * (world - 1) > 64:
        sink, peer = 1, 1
    assert len(owner) == sink + peer * (world - 1)
    assert owner.count(0) == sink
    for r in range(1, world):
        assert owner.count(r) == peer
    if sink > 1:          # the sink's slots are spread over the period, not bunched
        pos = [j for j, o in enumerate(owner) if o == 0]
        gaps = [b - a for a, b in zip(pos, pos[1:])]
        assert max(gaps) - min(gaps) <= 1
