"""Search for the shared-memory layout of the P4 fp64 regular-brick kernel (LayoutP4D in
wfx_stiffness.cu): brick lattice strides (Sx, Sy) modulo 16, the lane -> line maps of the three
roles and the tile row offsets such that every shared-memory access of a cell is free of bank
conflicts for 64-bit words (a warp's 25 active lanes are served as two half-warps, lanes 0-15 and
16-24, over 16 banks of 8 bytes).  One access pattern cannot be made conflict-free together with
the others (no stride triple admits it, see `pair_cost`): role I's read of its input line from the
staged dofs costs 3 wavefronts instead of 2.

Notation: a(q) = ascending lattice position of 1-D dof q in the [0, 1, interior] ordering.
  role K lane -> (i, j): dofs at  a(i)*Sx + a(j)*Sy + a(k),  tile columns A[k*PS + 5*a(i) + eA(a(j))],
                                                              AT[k*PS + 5*a(j) + eT(a(i))]
  role J lane -> (k, i): dofs at  a(i)*Sx + a(k) + a(m)*Sy,   tile row A[k*PS + 5*a(i) + ...]
  role I lane -> (k, j): dofs at  a(j)*Sy + a(k) + a(m)*Sx,   tile row AT[k*PS + 5*a(j) + ...]
Prints the tables as C initialisers.  Pure Python, runs in seconds."""
import itertools
from collections import Counter, defaultdict

N, S, T, PS = 5, 5, 2, 33   # Sx % 16, Sy % 16, tile plane stride
apos = [0, 4, 1, 2, 3]
ainv = [0, 2, 3, 4, 1]      # position -> 1-D dof index
perms = list(itertools.permutations(range(N)))
cells = [(a, b) for a in range(N) for b in range(N)]


def distinct(lanes, f):
    banks = [f(*c) % 16 for c in lanes]
    return len(set(banks)) == len(banks)


def cost(lanes, f):
    return max(Counter(f(*c) % 16 for c in lanes).values())


def partitions(f):
    """All splits of the 25 (alpha, beta) cells into 16 + 9 with distinct banks under f."""
    by = defaultdict(list)
    for c in cells:
        by[f(*c) % 16].append(c)
    if len(by) != 16 or max(len(v) for v in by.values()) > 2:
        return
    dbl = [k for k, v in by.items() if len(v) == 2]
    sgl = [k for k, v in by.items() if len(v) == 1]
    for mask in range(1 << len(dbl)):
        h0 = [by[k][0] for k in sgl] + [by[k][(mask >> n) & 1] for n, k in enumerate(dbl)]
        h1 = [by[k][1 - ((mask >> n) & 1)] for n, k in enumerate(dbl)]
        yield h0, h1


def search():
    # role K: cells are (alpha, beta) = (a(i), a(j))
    for h0, h1 in partitions(lambda a, b: S * a + T * b):
        for eA in perms:
            fA = lambda a, b: 5 * a + eA[b]
            if not (distinct(h0, fA) and distinct(h1, fA)):
                continue
            for eT in perms:
                fT = lambda a, b: 5 * b + eT[a]
                if distinct(h0, fT) and distinct(h1, fT):
                    return h0, h1, eA, eT
    raise SystemExit("no layout found")


def search_I(eT):
    # role I: cells are (alpha, gamma) = (a(j), a(k)); tile row bank PS*k + 5*a(j); dofs T*a(j) + a(k)
    best = None
    for h0, h1 in partitions(lambda a, g: PS * ainv[g] + 5 * a):
        c = cost(h0, lambda a, g: T * a + g) + cost(h1, lambda a, g: T * a + g)
        if best is None or c < best[0]:
            best = (c, h0, h1)
    return best


def main():
    h0, h1, eA, eT = search()
    laneK = h0 + h1
    # role J: cells (alpha, gamma) = (a(i), a(k)); dofs S*a(i) + a(k): lane = 5*alpha + gamma is the identity
    laneJ = [(l // N, l % N) for l in range(N * N)]
    fJ_tile = lambda a, g: PS * ainv[g] + 5 * a
    assert distinct(laneJ[:16], lambda a, g: S * a + g) and distinct(laneJ[16:], lambda a, g: S * a + g)
    assert distinct(laneJ[:16], fJ_tile) and distinct(laneJ[16:], fJ_tile)
    cI, i0, i1 = search_I(eT)
    laneI = i0 + i1
    print(f"// Sx % 16 = {S}, Sy % 16 = {T}, plane stride {PS}; role I dof read: {cI} wavefronts (ideal 2)")
    print("// role K lane -> i, j")
    print("{" + ", ".join(str(ainv[a]) for a, b in laneK) + "},")
    print("{" + ", ".join(str(ainv[b]) for a, b in laneK) + "},")
    print("// role J lane -> k, i")
    print("{" + ", ".join(str(ainv[g]) for a, g in laneJ) + "},")
    print("{" + ", ".join(str(ainv[a]) for a, g in laneJ) + "},")
    print("// role I lane -> k, j")
    print("{" + ", ".join(str(ainv[g]) for a, g in laneI) + "},")
    print("{" + ", ".join(str(ainv[a]) for a, g in laneI) + "},")
    print("// element offsets inside a row of A (by j) and of AT (by i)")
    print("{" + ", ".join(str(eA[apos[j]]) for j in range(N)) + "},")
    print("{" + ", ".join(str(eT[apos[i]]) for i in range(N)) + "},")


if __name__ == "__main__":
    main()
