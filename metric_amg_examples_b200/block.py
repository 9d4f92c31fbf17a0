"""The slice of cbc.block / FEniCS_ii that the reference's hot path touches, so that the
drivers' lines read the same without dolfin:

    block_vec / block_mat                 what get_system() returns (src/bidomain_2d.py:96-99)
    ii_convert(AA), ii_convert(bb)        block -> monolithic (src/bidomain_2d.py:178-179)
    ReductionOperator([len(W)], W)        block_vec <-> monolithic vector (src/utils.py:49)
    block_base, composition R.T*Minv*R    operator protocol of cbc.block (src/utils.py:53)

Vectors are numpy float64 arrays or torch CUDA float64 tensors; a block_vec is a list of them.
"""
import numpy as np
import scipy.sparse as sp


def _is_torch(x):
    return hasattr(x, "is_cuda")


class block_vec(list):
    """List of per-field vectors (cbc.block block_vec)."""

    def copy(self):
        return block_vec([v.clone() if _is_torch(v) else v.copy() for v in self])

    def norm(self):
        return float(np.sqrt(sum(float((v * v).sum()) for v in self)))

    def inner(self, other):
        return sum(float((a * b).sum()) for a, b in zip(self, other))

    def __add__(self, other):
        return block_vec([a + b for a, b in zip(self, other)])

    def __sub__(self, other):
        return block_vec([a - b for a, b in zip(self, other)])

    def __rmul__(self, s):
        return block_vec([s * a for a in self])

    def __mul__(self, s):
        return block_vec([a * s for a in self])

    def __neg__(self):
        return block_vec([-a for a in self])


class block_base:
    """cbc.block operator protocol: matvec, transpmult, __mul__ (vector -> apply,
    operator -> composition), create_vec, .T"""

    def matvec(self, b):
        raise NotImplementedError

    def transpmult(self, b):
        return self.matvec(b)

    def create_vec(self, dim=1):
        raise NotImplementedError

    @property
    def T(self):
        return block_transpose(self)

    def __mul__(self, other):
        if isinstance(other, block_base):
            return block_mul(self, other)
        return self.matvec(other)

    def __rmul__(self, scalar):
        return block_scale(self, scalar)


class block_transpose(block_base):
    def __init__(self, A):
        self.A = A

    def matvec(self, b):
        return self.A.transpmult(b)

    def transpmult(self, b):
        return self.A.matvec(b)

    def create_vec(self, dim=1):
        return self.A.create_vec(1 - dim)


class block_mul(block_base):
    """A*B*...: applied right to left (cbc.block block_mul)."""

    def __init__(self, *ops):
        self.chain = []
        for op in ops:
            self.chain.extend(op.chain if isinstance(op, block_mul) else [op])

    def _reduced_precond(self):
        """Minv if this is R.T * Minv * R with a device preconditioner (src/utils.py:53), else None."""
        if len(self.chain) != 3:
            return None
        Rt, M, R = self.chain
        if (hasattr(M, "hierarchy") and isinstance(R, ReductionOperator) and isinstance(Rt, block_transpose)
                and Rt.A is R):
            return M
        return None

    def matvec(self, b):
        M = self._reduced_precond()
        if M is not None and isinstance(b, block_vec) and len(b) == len(self.chain[2].sizes):
            # the block variant of the preconditioner: the blocks go to the device as they are and are
            # addressed by offsets there (mamg_apply_blocks) -- no concatenate / split copies
            M._ensure_device()
            return block_vec(M.hierarchy.apply_blocks(list(b)))
        for op in reversed(self.chain):
            b = op.matvec(b) if isinstance(op, block_base) else op * b
        return b

    def transpmult(self, b):
        for op in self.chain:
            b = op.transpmult(b)
        return b

    def create_vec(self, dim=1):
        return self.chain[0].create_vec(dim) if dim == 0 else self.chain[-1].create_vec(dim)


class block_scale(block_base):
    def __init__(self, A, s):
        self.A, self.s = A, s

    def matvec(self, b):
        return self.s * self.A.matvec(b)

    def create_vec(self, dim=1):
        return self.A.create_vec(dim)


class block_mat(block_base):
    """n x n array of scipy sparse blocks (what ii_assemble + apply_bc return)."""

    def __init__(self, blocks):
        self.blocks = np.empty((len(blocks), len(blocks[0])), dtype=object)
        for i, row in enumerate(blocks):
            for j, blk in enumerate(row):
                self.blocks[i, j] = blk

    def __getitem__(self, ij):
        return self.blocks[ij]

    def matvec(self, b):
        out = []
        for i in range(self.blocks.shape[0]):
            acc = None
            for j in range(self.blocks.shape[1]):
                blk = self.blocks[i, j]
                if blk is None:
                    continue
                t = blk @ b[j]
                acc = t if acc is None else acc + t
            out.append(acc)
        return block_vec(out)

    def create_vec(self, dim=1):
        n = [self.blocks[i, i].shape[dim] for i in range(self.blocks.shape[0])]
        return block_vec([np.zeros(k) for k in n])


class block_diag_mat(block_base):
    """xii.block_diag_mat([op_0, op_1, ...]): block i of the vector goes through operator i
    (src/utils.py:12: the block-LU `diag` preconditioner)."""

    def __init__(self, ops):
        self.ops = list(ops)

    def matvec(self, b):
        if len(b) != len(self.ops):
            raise ValueError(f"block_diag_mat of {len(self.ops)} blocks applied to a {len(b)}-block vector")
        return block_vec([op * v if not isinstance(op, block_base) else op.matvec(v) for op, v in zip(self.ops, b)])

    def transpmult(self, b):
        return block_vec([op.transpmult(v) if isinstance(op, block_base) else op * v for op, v in zip(self.ops, b)])

    def create_vec(self, dim=1):
        return block_vec([op.create_vec(dim) for op in self.ops])


def ii_convert(obj):
    """Collapse a block_mat to one monolithic CSR matrix / a block_vec to one vector, blocks
    concatenated in order (xii.ii_convert, src/bidomain_2d.py:178-179)."""
    if isinstance(obj, block_mat):
        A = sp.bmat([[obj.blocks[i, j] for j in range(obj.blocks.shape[1])]
                     for i in range(obj.blocks.shape[0])], format="csr")
        A.sort_indices()
        return A
    if isinstance(obj, block_vec):
        if _is_torch(obj[0]):
            import torch
            return torch.cat(list(obj))
        return np.concatenate(list(obj))
    return obj


def split_blocks(A, sizes):
    """Inverse of ii_convert for matrices: monolithic CSR -> block_mat with the given sizes."""
    offs = np.concatenate([[0], np.cumsum(sizes)])
    A = A.tocsr()
    return block_mat([[A[offs[i]:offs[i + 1], offs[j]:offs[j + 1]].tocsr() for j in range(len(sizes))]
                      for i in range(len(sizes))])


class ReductionOperator(block_base):
    """xii.ReductionOperator([len(W)], W): maps the len(W)-block block_vec to the 1-block
    monolithic vector (concatenation); .T splits it back (src/utils.py:49,53)."""

    def __init__(self, offsets, W):
        assert list(offsets) == [len(W)], "only the full reduction used by the reference is supported"
        self.sizes = [w.dim() if hasattr(w, "dim") else int(w) for w in W]
        self.offs = np.concatenate([[0], np.cumsum(self.sizes)])

    def matvec(self, b):
        return ii_convert(block_vec(b))

    def transpmult(self, x):
        if isinstance(x, block_vec) and len(x) == 1:
            x = x[0]
        return block_vec([x[self.offs[i]:self.offs[i + 1]] for i in range(len(self.sizes))])

    def create_vec(self, dim=1):
        if dim == 1:
            return block_vec([np.zeros(k) for k in self.sizes])
        return np.zeros(self.offs[-1])
