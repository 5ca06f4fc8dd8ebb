"""NumPy model of the 2-D block-cyclic Cholesky / solve of g3py_b200/csrc/dist.cu, one instance per rank.

Test infrastructure only: it restates the C++ host schedule (pieces, piece-major panel buffers, diagonal-block exchange
inside a process column, grouped piece broadcasts, per-panel all-reduce of the substitution) with NumPy blocks, so the
index arithmetic and the message pattern can be exercised on CPUs over gloo (world 2 and 4) and compared with the
layout the library reports (g3_dist_layout)."""
import numpy as np


def first_blk(J, p, Pr):
    return J + ((p - J % Pr) % Pr + Pr) % Pr


def cnt_blk(J, p, Pr, nP):
    f = first_blk(J, p, Pr)
    return (nP - 1 - f) // Pr + 1 if f < nP else 0


class LocalComm:
    """world == 1"""
    rank, world = 0, 1

    def bcast(self, a, root):
        return a

    def sendrecv_diag(self, a, root, peers, me):
        return a

    def allreduce(self, a):
        return a


class GlooComm:
    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def bcast(self, a, root):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a))
        self.dist.broadcast(t, src=root)
        return t.numpy()

    def sendrecv_diag(self, a, root, peers, me):
        """root sends `a` to every rank in peers; peers receive into a buffer of the same shape"""
        import torch
        if me == root:
            for r in peers:
                self.dist.send(torch.from_numpy(np.ascontiguousarray(a)), dst=r)
            return a
        t = torch.empty(a.shape, dtype=torch.float64)
        self.dist.recv(t, src=root)
        return t.numpy()

    def allreduce(self, a):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a))
        self.dist.all_reduce(t)
        return t.numpy()


class DistModel:
    def __init__(self, K, nb, Pr, Pc, comm):
        self.N = K.shape[0]
        self.nb, self.Pr, self.Pc, self.comm = nb, Pr, Pc, comm
        self.nP = self.N // nb
        self.p, self.q = comm.rank % Pr, comm.rank // Pr
        self.store = {}
        for J in range(self.q, self.nP, Pc):
            cnt = cnt_blk(J, self.p, Pr, self.nP)
            if cnt == 0:
                continue
            I0 = first_blk(J, self.p, Pr)
            rows = np.concatenate([np.arange((I0 + i * Pr) * nb, (I0 + i * Pr + 1) * nb) for i in range(cnt)])
            self.store[J] = K[rows][:, J * nb:(J + 1) * nb].copy()
        self.logdet = 0.0

    def rank_of(self, p, q):
        return q * self.Pr + p

    def _panel(self, J):
        """every rank ends up with all pieces of panel J (piece-major list)"""
        qJ = J % self.Pc
        pieces = []
        for pp in range(self.Pr):
            c = cnt_blk(J, pp, self.Pr, self.nP)
            if c == 0:
                pieces.append(None)
                continue
            mine = self.q == qJ and pp == self.p
            buf = self.store[J] if mine else np.empty((c * self.nb, self.nb))
            pieces.append(self.comm.bcast(buf, self.rank_of(pp, qJ)))
        return pieces

    def _block(self, pieces, K, J):
        pp = K % self.Pr
        idx = (K - first_blk(J, pp, self.Pr)) // self.Pr
        return pieces[pp][idx * self.nb:(idx + 1) * self.nb]

    def factor(self):
        nb, Pr, Pc, nP, p, q = self.nb, self.Pr, self.Pc, self.nP, self.p, self.q
        for J in range(nP):
            qJ, pd = J % Pc, J % Pr
            in_col = q == qJ
            cnt = cnt_blk(J, p, Pr, nP)
            Ld = None
            if in_col and p == pd:
                L = np.linalg.cholesky(self.store[J][:nb])
                self.store[J][:nb] = L
                self.logdet += float(np.log(np.diag(L)).sum())
                Ld = L
            if Pr > 1 and in_col:
                peers = [self.rank_of(pp, q) for pp in range(Pr) if pp != pd and cnt_blk(J, pp, Pr, nP) > 0]
                if p == pd or cnt > 0:
                    Ld = self.comm.sendrecv_diag(Ld if Ld is not None else np.empty((nb, nb)), self.rank_of(pd, q), peers,
                                                 self.comm.rank)
            if in_col and cnt > 0:
                skip = 1 if p == pd else 0
                if cnt - skip > 0:
                    rest = self.store[J][skip * nb:]
                    rest[:] = np.linalg.solve(Ld, rest.T).T            # piece <- piece L_JJ^-T
            pieces = self._panel(J)
            for K in range(J + 1, nP):
                if K % Pc != q:
                    continue
                c = cnt_blk(K, p, Pr, nP)
                if c == 0:
                    continue
                IK = first_blk(K, p, Pr)
                a0 = (IK - first_blk(J, p, Pr)) // Pr
                A = pieces[p][a0 * nb:(a0 + c) * nb]
                self.store[K] -= A @ self._block(pieces, K, J).T
        return self

    def solve(self, delta):
        nb, Pr, Pc, nP, p, q = self.nb, self.Pr, self.Pc, self.nP, self.p, self.q
        c = delta.copy() if self.comm.rank == 0 else np.zeros(self.N)
        u = np.zeros(self.N)
        beta = 0.0
        for J in range(nP):
            qJ, pd = J % Pc, J % Pr
            seg = self.comm.allreduce(c[J * nb:(J + 1) * nb].copy())
            uJ = np.zeros(nb)
            if q == qJ and p == pd:
                uJ = np.linalg.solve(np.tril(self.store[J][:nb]), seg)
                beta += float(uJ @ uJ)
            uJ = self.comm.bcast(uJ, self.rank_of(pd, qJ))
            u[J * nb:(J + 1) * nb] = uJ
            cnt = cnt_blk(J, p, Pr, nP)
            if q == qJ and cnt > 0:
                skip = 1 if p == pd else 0
                I0 = first_blk(J, p, Pr)
                for i in range(skip, cnt):
                    I = I0 + i * Pr
                    c[I * nb:(I + 1) * nb] -= self.store[J][i * nb:(i + 1) * nb] @ uJ
        beta = float(self.comm.allreduce(np.array([beta]))[0])
        return u, beta

    def check_against(self, Lref, tol=1e-10):
        nb, Pr = self.nb, self.Pr
        for J, piece in self.store.items():
            I0 = first_blk(J, self.p, Pr)
            for i in range(piece.shape[0] // nb):
                I = I0 + i * Pr
                got = piece[i * nb:(i + 1) * nb]
                want = Lref[I * nb:(I + 1) * nb, J * nb:(J + 1) * nb]
                if I == J:
                    got = np.tril(got)
                assert np.abs(got - want).max() < tol, (J, I, np.abs(got - want).max())
