import random, heapq
def simulate(n_virt, n_st, total_tiles, n_cta, seed):
    """event-driven model of the dynamic tile scheduler protocol of fc_chain_kernel (one global counter, n_cta CTAs)."""
    rng = random.Random(seed)
    counter = 0                          # tiles handed out beyond the static n_st*n_cta
    processed_tiles = []
    ctas = []
    for c in range(n_cta):
        ring = {}                        # local index -> tile or None(invalid); presence = published
        for i in range(n_st):
            t = c + i * n_cta
            ring[i] = t if t < total_tiles else None
        ctas.append(dict(ring=ring, next=[q for q in range(n_virt)], alive=[True] * n_virt, busy=[False] * n_virt))
    def fetch():
        nonlocal counter
        t = n_st * n_cta + counter; counter += 1
        return t if t < total_tiles else None
    events = []  # (time, cta, slot, kind)
    time = 0.0
    # each slot repeatedly: wait until ring has its next index; if invalid: die + publish(i+n_st, invalid); else: L1 takes delay, then publish(i+n_st, fetch()), then rest of chain delay, then next
    pending = [(0.0, c, q) for c in range(n_cta) for q in range(n_virt)]
    heapq.heapify(pending)
    steps = 0
    blocked = {}
    while pending:
        steps += 1
        if steps > 2_000_000: return "runaway"
        t, c, q = heapq.heappop(pending)
        st = ctas[c]
        if not st["alive"][q]: continue
        i = st["next"][q]
        if i not in st["ring"]:
            blocked[(c, q)] = i           # wait for publication
            continue
        tile = st["ring"][i]
        def publish(j, val, when):
            assert j not in st["ring"], ("double publish", c, j)
            st["ring"][j] = val
            for (cc, qq), jj in list(blocked.items()):
                if cc == c and jj == j:
                    del blocked[(cc, qq)]
                    heapq.heappush(pending, (when, cc, qq))
        if tile is None:
            st["alive"][q] = False
            publish(i + n_st, None, t)
            continue
        processed_tiles.append(tile)
        t_l1 = t + rng.uniform(0.5, 2.0)
        publish(i + n_st, fetch(), t_l1)
        st["next"][q] = i + n_virt
        heapq.heappush(pending, (t_l1 + rng.uniform(1.0, 6.0), c, q))
    if blocked: return f"DEADLOCK {blocked}"
    if any(a for st in ctas for a in st["alive"]): return "alive slots left"
    if sorted(processed_tiles) != list(range(total_tiles)): return f"tiles wrong: {len(processed_tiles)} vs {total_tiles}, dup={len(processed_tiles)-len(set(processed_tiles))}"
    return "ok"
bad = 0
for n_virt in (1, 2, 3, 4, 6):
    for n_st in (2, 3, 4, 5, 6, 8):
        for total in (0, 1, 5, 17, 40, 200):
            for n_cta in (1, 3, 7):
                for seed in range(6):
                    r = simulate(n_virt, n_st, total, n_cta, seed)
                    if r != "ok":
                        bad += 1
                        if bad < 15: print(n_virt, n_st, total, n_cta, seed, r)
print("failures:", bad)
