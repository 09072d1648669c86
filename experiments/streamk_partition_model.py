"""Host model of the stream-K partition arithmetic of w4_streamk_gemv_UNVALIDATED.patch: every (row tile, k slot) unit
is covered exactly once, piece indices are in range, and the ticket count of a tile equals its number of partial
pieces.  Run: python experiments/streamk_partition_model.py"""


def plan(N, K, gs, sms=148):
    ntiles, ktiles = N // 16, K // 64
    tg = 2 if gs // 64 >= 2 else 1
    nslots = ktiles // tg
    slots = 2 * sms
    units = ntiles * nslots
    grid = slots
    min_range = (nslots + 1) // 2
    if units // grid < min_range:
        grid = units // min_range
    return ntiles, nslots, units, grid


def check(N, K, gs=128):
    ntiles, nslots, units, G = plan(N, K, gs)
    cover = [[0] * nslots for _ in range(ntiles)]
    arrivals = [0] * ntiles
    expected = {}
    max_pieces = 0
    for c in range(G):
        u0, u1 = c * units // G, (c + 1) * units // G
        assert u1 > u0, (N, K, c, "empty range")
        while u0 < u1:
            r = u0 // nslots
            a = u0 - r * nslots
            b = min(nslots, a + (u1 - u0))
            for s in range(a, b):
                cover[r][s] += 1
            whole = a == 0 and b == nslots
            c_first = ((r * nslots + 1) * G - 1) // units
            c_last = ((r * nslots + nslots) * G - 1) // units
            npieces, piece = c_last - c_first + 1, c - c_first
            assert 0 <= piece < npieces, (N, K, c, r, piece, npieces)
            max_pieces = max(max_pieces, npieces)
            if whole:
                assert npieces == 1
            else:
                arrivals[r] += 1
                expected[r] = npieces
            u0 += b - a
    assert all(v == 1 for row in cover for v in row), "coverage"
    for r, n in expected.items():
        assert arrivals[r] == n, (N, K, r, arrivals[r], n)
    return G, max_pieces


if __name__ == "__main__":
    for N, K in [(3584, 3584), (3584, 18944), (2048, 4096), (3584, 1792), (1792, 9472), (896, 4736), (3584, 4736)]:
        print(N, K, "grid, max pieces =", check(N, K))
