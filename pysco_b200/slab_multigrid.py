"""multigrid.linear (multigrid.py:23-83) on x-slabs: the host sequencing of the V-cycle over the ghost-plane kernels of
csrc/slab_mg.cu, as SURVEY 8(e) lays it out.

* Every level keeps the decomposition of the fine grid: rank r owns planes [r n/P, (r+1) n/P) of the n^3 level.
  Potential-like arrays carry one ghost plane on each side along x; ``exchange_planes`` refreshes them before every
  half-sweep of the red-black smoother (the two colours of a sweep read each other across the slab boundary), before a
  residual and before a prolongation.
* Restriction and prolongation are local: the two children planes of a coarse plane live on the same rank.
* Once a coarse level would have fewer than two planes per rank it is gathered on every rank and solved redundantly
  with the single-domain kernels (csrc/multigrid.cu), which costs one all-gather of at most (2P)^3 cells.
* Norms (residual, truncation error) are local double sums, all-reduced.

The cycle is the reference's: Npre smoothing sweeps, R(b - Lx), first guess -h^2/6 b, recursion (or Npre sweeps on the
coarsest 4^3 level), x += P(correction), Npost sweeps; stopping rule and truncation-error refresh as multigrid.py:62-80.

The f(R) scalaron uses the same decomposition for the full-approximation-storage cycle (multigrid.FAS,
multigrid.py:88-140, 521-579): nonlinear red-black sweeps (cubic / quartic closed-form roots), coarse problem
L(x_c) = 4 R(residual) + L(R x), correction x_c - R x prolongated back.
"""
import logging

import numpy as np
import torch

from . import _lib

F_RELAX = np.float32(1.25)   # laplacian.py:1053, cubic.py:1094, 1133


class SlabMultigrid:
    def __init__(self, comm, ops, ncells_1d):
        self.comm, self.ops, self.N = comm, ops, int(ncells_1d)
        self.P, self.rank = comm.size, comm.rank
        nxl = self.N // self.P
        if self.N % self.P or nxl < 2 or (nxl & (nxl - 1)):
            raise ValueError(f"slab multigrid needs a power-of-two number of planes per rank (>= 2), got "
                             f"{self.N} / {self.P}")
        self.nxl = nxl

    # -- ghosts
    def exchange(self, xg, nxl):
        """ghost plane 0 <- the left neighbour's last owned plane, ghost plane nxl + 1 <- the right neighbour's first"""
        from_left, from_right = self.comm.exchange_planes(xg[1:2], xg[nxl:nxl + 1])
        xg[0].copy_(from_left[0])
        xg[nxl + 1].copy_(from_right[0])

    # -- laplacian.smoothing (laplacian.py:1026-1055): red (odd i+j+k) then black, ghosts refreshed before each
    def smoothing(self, xg, b, n, nxl, sweeps):
        x0 = self.rank * nxl
        for _ in range(int(sweeps)):
            for colour in (1, 0):
                self.exchange(xg, nxl)
                self.ops.mg_gs_colour(xg, b, nxl, n, x0, colour, F_RELAX)

    def _norm(self, sumsq):
        self.comm.allreduce_sum_(sumsq)
        return np.float32(np.sqrt(sumsq.item()))

    def residual_error(self, xg, b, n, nxl):
        """laplacian.residual_error (laplacian.py:327-381)"""
        self.exchange(xg, nxl)
        Lx = self.ops.mg_operator(xg, nxl, n)
        return self._norm(self.ops.mg_diff_sumsq(b, 1.0, Lx))

    def truncation_error(self, xg, n, nxl):
        """laplacian.truncation_error (laplacian.py:502-533): || R(L x) - L(R x) ||"""
        ops = self.ops
        nc, nxlc = n // 2, nxl // 2
        self.exchange(xg, nxl)
        RLx = ops.mg_restriction(ops.mg_operator(xg, nxl, n), nxl, n, 1.0)
        Rxg = torch.empty((nxlc + 2, nc, nc), dtype=torch.float32, device=xg.device)
        ops.mg_restriction(xg[1:nxl + 1], nxl, n, 1.0, out=Rxg[1:nxlc + 1])
        self.exchange(Rxg, nxlc)
        LRx = ops.mg_operator(Rxg, nxlc, nc)
        return self._norm(ops.mg_diff_sumsq(RLx, 1.0, LRx))

    @staticmethod
    def first_guess_factor(n):
        """laplacian.initialise_potential (laplacian.py:765-796): x = -h^2/6 b, the factor rounded like the kernel's"""
        h = float(np.float32(1.0) / np.float32(n))
        return np.float32(-h * h / 6.0)

    def _gather_level(self, planes, nc):
        """all ranks' owned planes of a coarse level -> the full [nc, nc, nc] cube on every rank"""
        P = self.P
        send = planes.reshape(1, -1).expand(P, -1).contiguous()
        out = torch.empty_like(send)
        self.comm.all_to_all_equal(send, out)
        return out.view(nc, nc, nc)

    # -- multigrid.V_cycle (multigrid.py:474-517)
    def v_cycle(self, xg, b, n, nxl, param, nlevel=0):
        ops = self.ops
        nc, nxlc = n // 2, nxl // 2
        self.smoothing(xg, b, n, nxl, param["Npre"])
        self.exchange(xg, nxl)
        res_c = ops.mg_restrict_residual(xg, b, nxl, n)              # [nxlc, nc, nc]
        coarsest = nlevel >= (param["ncoarse"] - 3)
        if nxlc >= 2:
            cg = torch.empty((nxlc + 2, nc, nc), dtype=torch.float32, device=xg.device)
            cg[1:nxlc + 1].copy_(res_c)
            ops.affine(cg[1:nxlc + 1], self.first_guess_factor(nc), 0.0)
            if coarsest:
                self.smoothing(cg, res_c, nc, nxlc, param["Npre"])
            else:
                self.v_cycle(cg, res_c, nc, nxlc, param, nlevel + 1)
            self.exchange(cg, nxlc)
        else:
            # one coarse plane per rank: gather the level and solve it on every rank with the single-domain kernels
            cube = self._gather_level(res_c, nc)
            corr = ops.mg_cube_solve(cube, param, nlevel, coarsest)
            c0 = self.rank * nxlc
            idx = torch.tensor([(c0 - 1 + p) % nc for p in range(nxlc + 2)], dtype=torch.int64, device=corr.device)
            cg = corr.index_select(0, idx)
        ops.mg_add_prolongation(xg, cg, nxlc, nc)
        self.smoothing(xg, b, n, nxl, param["Npost"])

    # -- multigrid.linear (multigrid.py:23-83)
    def linear(self, xg, b, param):
        """xg [nxl + 2, N, N]: first guess in the owned planes, solution on return (ghost planes are scratch);
        b [nxl, N, N]."""
        if param.get("compute_additional_field", False) and "fr" == param["theory"].casefold():
            raise ValueError("Linear should not be used for scalaron field")
        n, nxl = self.N, self.nxl
        # the second solve of QUMOND keeps its own tolerance (multigrid.py:58-72)
        mond_pass = (not param.get("compute_additional_field", False)) and "mond" == param["theory"].casefold()
        if ("tolerance" not in param) or (param["nsteps"] % 3) == 0:
            logging.info("Compute Truncation error")
            param["tolerance_mond" if mond_pass else "tolerance"] = param["epsrel"] * self.truncation_error(xg, n, nxl)
        tolerance = param["tolerance_mond"] if mond_pass else param["tolerance"]
        logging.info("Start linear Multigrid (slab)")
        residual_err = 1e30
        while residual_err > tolerance:
            self.v_cycle(xg, b, n, nxl, param, 0)
            residual_error_tmp = self.residual_error(xg, b, n, nxl)
            logging.info(f"{residual_error_tmp=} {tolerance=}")
            if residual_error_tmp < tolerance or residual_err / residual_error_tmp < 2:
                break
            residual_err = residual_error_tmp
        return xg

    # ------------------------------------------------------------------------------------ f(R): FAS
    @staticmethod
    def fr_kind(param):
        n = param["fR_n"]
        if n == 1:
            return _lib.OP_CUBIC
        if n == 2:
            return _lib.OP_QUARTIC
        raise NotImplementedError(f"Only f(R) with n = 1 and 2, currently {param['fR_n']=}")

    def smoothing_fr(self, xg, b, rhs, q, kind, n, nxl, sweeps):
        """cubic / quartic.smoothing[_with_rhs] (cubic.py:1064-1140); rhs = None on the finest level"""
        x0 = self.rank * nxl
        for _ in range(int(sweeps)):
            for colour in (1, 0):
                self.exchange(xg, nxl)
                self.ops.mg_gs_colour_fr(xg, b, rhs, q, nxl, n, x0, colour, F_RELAX, kind)

    def residual_error_fr(self, xg, b, q, kind, n, nxl):
        """cubic / quartic.residual_error (cubic.py:844-901): sqrt(sum L(x)^2)"""
        self.exchange(xg, nxl)
        Lx = self.ops.mg_operator_fr(xg, b, q, nxl, n, kind)
        return self._norm(self.ops.mg_diff_sumsq(Lx, 2.0, Lx))     # (2 L - L)^2 = L^2, exactly

    def truncation_error_fr(self, xg, b, q, kind, n, nxl):
        """cubic / quartic.truncation_error (cubic.py:1021-1061): || 4 R(L x) - L(R x; R b) ||"""
        ops = self.ops
        nc, nxlc = n // 2, nxl // 2
        self.exchange(xg, nxl)
        RLx = ops.mg_restriction(ops.mg_operator_fr(xg, b, q, nxl, n, kind), nxl, n, 1.0)
        Rxg = torch.empty((nxlc + 2, nc, nc), dtype=torch.float32, device=xg.device)
        ops.mg_restriction(xg[1:nxl + 1], nxl, n, 1.0, out=Rxg[1:nxlc + 1])
        self.exchange(Rxg, nxlc)
        LRx = ops.mg_operator_fr(Rxg, ops.mg_restriction(b, nxl, n, 1.0), q, nxlc, nc, kind)
        return self._norm(ops.mg_diff_sumsq(RLx, 4.0, LRx))

    # -- multigrid.V_cycle_FAS (multigrid.py:521-579)
    def v_cycle_fas(self, xg, b, n, nxl, param, nlevel, rhs, q, kind):
        ops = self.ops
        nc, nxlc = n // 2, nxl // 2
        self.smoothing_fr(xg, b, rhs, q, kind, n, nxl, param["Npre"])
        b_c = ops.mg_restriction(b, nxl, n, 1.0)
        self.exchange(xg, nxl)
        Lx = ops.mg_operator_fr(xg, b, q, nxl, n, kind)
        if rhs is None:
            res_c = ops.mg_restriction(Lx, nxl, n, -1.0)            # minus_restriction(L x)
        else:
            ops.lincomb(Lx, -1.0, rhs, 1.0)                         # rhs - L x
            res_c = ops.mg_restriction(Lx, nxl, n, 1.0)
        del Lx
        coarsest = nlevel >= (param["ncoarse"] - 3)
        if nxlc >= 2:
            cg = torch.empty((nxlc + 2, nc, nc), dtype=torch.float32, device=xg.device)
            own = cg[1:nxlc + 1]
            ops.mg_restriction(xg[1:nxl + 1], nxl, n, 1.0, out=own)  # x_c = R x, the first guess of the coarse problem
            x_c = own.clone()
            self.exchange(cg, nxlc)
            L_c = ops.mg_operator_fr(cg, b_c, q, nxlc, nc, kind)
            ops.lincomb(res_c, 4.0, L_c, 1.0)                       # coarse right-hand side 4 R(res) + L(R x)
            del L_c
            if coarsest:
                self.smoothing_fr(cg, b_c, res_c, q, kind, nc, nxlc, param["Npre"])
            else:
                self.v_cycle_fas(cg, b_c, nc, nxlc, param, nlevel + 1, res_c, q, kind)
            ops.axpy(own, x_c, -1.0)                                # correction = x_c - R x
            self.exchange(cg, nxlc)
        else:
            # one coarse plane per rank: the coarse problem is gathered and solved on every rank
            x_c = ops.mg_restriction(xg[1:nxl + 1], nxl, n, 1.0)
            corr = ops.mg_cube_solve_fas(self._gather_level(x_c, nc), self._gather_level(b_c, nc),
                                         self._gather_level(res_c, nc), param, nlevel, coarsest)
            c0 = self.rank * nxlc
            idx = torch.tensor([(c0 - 1 + p) % nc for p in range(nxlc + 2)], dtype=torch.int64, device=corr.device)
            cg = corr.index_select(0, idx)
        ops.mg_add_prolongation(xg, cg, nxlc, nc)
        self.smoothing_fr(xg, b, rhs, q, kind, n, nxl, param["Npost"])

    # -- multigrid.FAS (multigrid.py:88-140)
    def fas(self, xg, b, param):
        """xg [nxl + 2, N, N]: first guess of the scalaron in the owned planes, solution on return; b [nxl, N, N] the
        density term; q = param["fR_q"]."""
        n, nxl = self.N, self.nxl
        kind = self.fr_kind(param)
        q = np.float32(param["fR_q"])
        if ("tolerance_FAS" not in param) or (param["nsteps"] % 3) == 0:
            logging.info("Compute FAS Truncation error")
            param["tolerance_FAS"] = param["epsrel"] * self.truncation_error_fr(xg, b, q, kind, n, nxl)
        tolerance = param["tolerance_FAS"]
        logging.info("Start Full-Approximation Storage Multigrid (slab)")
        residual_err = 1e30
        while residual_err > tolerance:
            self.v_cycle_fas(xg, b, n, nxl, param, 0, None, q, kind)
            residual_error_tmp = self.residual_error_fr(xg, b, q, kind, n, nxl)
            logging.info(f"{residual_error_tmp=} {tolerance=}")
            if residual_error_tmp < tolerance or residual_err / residual_error_tmp < 2:
                break
            residual_err = residual_error_tmp
        return xg
