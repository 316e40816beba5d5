"""CPU: the oracle's PRM::solve restatement (oracle/oxg_oracle.cpp, orc_prm_solve — prm.rs:227-307)
against an independent pure-Python walk of the same reference lines on the reference's own wall
fixture (oxmpl/tests/prm_rvss_tests.rs: PRM::new(5.0, 0.5), start (1,5), goal (9,5) r 0.5)."""
from collections import deque

import numpy as np


def _python_prm_solve(orc, ow, sp_o, rows, off, idx, radius, res, start, goal, gr):
    n = len(rows)
    if n == 0:
        return 3, []
    if not orc.is_valid(ow, sp_o, np.asarray([start]))[0]:
        return 2, []
    edges = [[] for _ in range(n)]
    for i in range(n):                                           # prm.rs:135, 143-145
        mine = [int(j) for j in idx[int(off[i]):int(off[i + 1])]]
        edges[i].extend(mine)
        for j in mine:
            edges[j].append(i)
    d = np.sqrt(((np.asarray(start) - rows) ** 2).sum(1))        # close enough to pick candidates
    cand = np.nonzero(d < radius * 1.0001)[0]
    conn = []
    for i in cand:                                               # prm.rs:249-256
        di = float(np.sqrt((start[0] - rows[i, 0]) * (start[0] - rows[i, 0]) +
                           (start[1] - rows[i, 1]) * (start[1] - rows[i, 1])))
        if di < radius and orc.check_motion(ow, sp_o, np.asarray([start]), rows[i:i + 1], res)[0][0]:
            conn.append(int(i))
    is_goal = [float(np.sqrt((r[0] - goal[0]) * (r[0] - goal[0]) + (r[1] - goal[1]) * (r[1] - goal[1]))) <= gr
               for r in rows]                                    # prm.rs:259-264
    if not conn or not any(is_goal):
        return 1, []
    queue = deque(conn)                                          # prm.rs:271 (seeded with a clone ...)
    parent, visited = {}, [False] * n
    for i in conn:                                               # ... and pushed again, prm.rs:275-279
        queue.append(i)
        parent[i] = None
        visited[i] = True
    while queue:
        cur = queue.popleft()
        if is_goal[cur]:
            out = []
            while cur is not None:
                out.append(cur)
                cur = parent[cur]
            return 0, out[::-1]
        for j in edges[cur]:
            if not visited[j]:
                visited[j] = True
                parent[j] = cur
                queue.append(j)
    return 1, []


def test_oracle_prm_solve_follows_the_reference_walk(orc):
    wall = np.array([[4.75, 2.0, 5.25, 8.0]])
    sp_o = orc.space(2)
    ow, keep = orc.world(2, aabbs=wall)
    res = orc.resolution(orc.rv_max_extent([0.0, 0.0], [10.0, 10.0]))
    rng = np.random.default_rng(7)
    samples = rng.random((2500, 2)) * 10.0
    rows, off, idx, _, _ = orc.prm_radius(ow, sp_o, 0.5, res, samples)
    assert len(rows) > 2000 and off[-1] > 0
    cases = [((1.0, 5.0), (9.0, 5.0), 0.5), ((5.0, 5.0), (9.0, 5.0), 0.5), ((1.0, 5.0), (9.0, 5.0), 1e-9),
             ((9.5, 9.5), (0.5, 0.5), 0.5), ((2.0, 9.0), (8.0, 1.0), 0.3)]
    seen = set()
    for start, goal, gr in cases:
        status, path = orc.prm_solve(ow, sp_o, rows, off, idx, 0.5, res, start, goal, gr)
        want_status, want_path = _python_prm_solve(orc, ow, sp_o, rows, off, idx, 0.5, res, start, goal, gr)
        assert status == want_status
        assert list(path) == want_path
        seen.add(status)
    assert seen == {0, 1, 2}
    status, path = orc.prm_solve(ow, sp_o, np.zeros((0, 2)), np.zeros(1, np.uint64), np.zeros(0, np.uint32), 0.5,
                                 res, (1.0, 5.0), (9.0, 5.0), 0.5)
    assert status == 3 and len(path) == 0
