"""Functional model of the CUDA systolic projection schedule (tfs_systolic.cuh), in NumPy.

TEST INFRASTRUCTURE ONLY.  This file is not an oracle of the reference: it is an executable
specification of the *kernel's data flow* (which lane holds which operand at which step, what
crosses lanes / warps / bands, what is staged from and to memory), used on the CPU to prove
that the schedule reproduces the reference's lexicographic Gauss-Seidel sweep
(simulator.rs:157-189) bit-for-bit before the same index algebra is run on a GPU.

Schedule.  One pass fuses KB projection iterations.  A band owns LANES consecutive values of
the skewed coordinate q = i + k (column + iteration); virtual lane L of band b has
q = b*LANES + L - 1.  At step t the lane works, for every k in [0, KB), on

        column i = q - k,        row j = t - L - k - 1.

With this skew every operand of cell (i, j, k) was produced at step t-1 by the same lane or by
lane L-1 (u/v faces), or at step t-2 by lane L-1 (pressure, cell flags):

    uC (u[i,j]   after (i-1,j,k))   <- lane L-1, iteration k,   step t-1   ("uRp")
    uR (u[i+1,j] after iter k-1)    <- same lane, iteration k-1, step t-1   ("uCf")
    vC (v[i,j]   after (i,j-1,k))   <- same lane, iteration k,   step t-1   ("vTp")
    vT (v[i,j+1] after iter k-1)    <- lane L-1, iteration k-1, step t-1   ("vCf")
    p, flags of (i,j) after k-1     <- lane L-1, iteration k-1, step t-2

k = 0 operands come from memory (previous pass / input fields); after k = KB-1 the lane holds
the final u[i,j], v[i,j], p[i,j] of the pass and stores them.  Lane 0 of a band takes its
"lane L-1" operands from the record stream written by the last lane of band b-1.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def cell_update_vec(f4, omega, pc, uC, uR, vC, vT, p):
    """Vectorised simulator.rs:161-186 on float32 arrays; f4 = neighbour-fluid bits (0 => skip)."""
    f4 = f4.astype(np.int64)
    live = f4 != 0
    sL = ((f4 & 1) != 0).astype(F)
    sR = ((f4 & 2) != 0).astype(F)
    sB = ((f4 & 4) != 0).astype(F)
    sT = ((f4 & 8) != 0).astype(F)
    n = ((sR + sT) + sL) + sB
    n_safe = np.where(live, n, F(1))
    with np.errstate(all="ignore"):
        div = ((uR - uC) + vT) - vC
        corr = omega * ((-div) / n_safe)
        uC2 = uC - corr * sL
        uR2 = uR + corr * sR
        vC2 = vC - corr * sB
        vT2 = vT + corr * sT
        p2 = p + pc * corr
    return (np.where(live, uC2, uC), np.where(live, uR2, uR), np.where(live, vC2, vC),
            np.where(live, vT2, vT), np.where(live, p2, p))


def systolic_pass(u, v, p, flags, W, H, KB, k_act, LANES, omega, pc, gdt, first_pass):
    """One temporal pass over the whole grid; returns new (u, v, p).  u, v, p, flags are flat
    column-major arrays (idx = x*H + y).  Bands are run one after the other, each leaving the
    record stream its right neighbour consumes."""
    omega, pc, gdt = F(omega), F(pc), F(gdt)
    n_bands = (W + KB - 1 + 1 + LANES - 1) // LANES      # q ranges over [-1, W + KB - 2]
    T = H + LANES + KB                                   # steps per band (rows H.. are phantom)
    uo, vo, po = np.zeros_like(u), np.zeros_like(v), np.zeros_like(p)
    u2, v2, p2, f2 = (a.reshape(W, H) for a in (u, v, p, flags))
    uo2, vo2, po2 = (a.reshape(W, H) for a in (uo, vo, po))

    def ld(arr, col, row, dflt=0):
        ok = (col >= 0) & (col < W) & (row >= 0) & (row < H)
        out = np.full(col.shape, dflt, arr.dtype)
        out[ok] = arr[col[ok], row[ok]]
        return out

    prev_rec = None                                      # records of band b-1: list over steps
    for b in range(n_bands):
        L = np.arange(LANES)
        q = b * LANES + L - 1
        uL = np.zeros((KB, LANES), F); uRo = np.zeros((KB, LANES), F)
        vB = np.zeros((KB, LANES), F); vTo = np.zeros((KB, LANES), F)
        pin = np.zeros((KB, LANES), F); pd = np.zeros((KB, LANES), F)
        fl = np.zeros((KB, LANES), np.uint8); fld = np.zeros((KB, LANES), np.uint8)
        rec = []

        def prep_k0(tau):
            """operands of iteration 0 for step tau, from memory (the staged tiles in CUDA)"""
            row = tau - L - 1
            uRo[0] = ld(u2, q + 1, row)
            vv = ld(v2, q, row + 1)
            if first_pass:                               # add_gravity fused (simulator.rs:137-149)
                live_above = (ld(f2, q, row + 1) & 16) != 0
                vv = np.where(live_above, vv + gdt, vv)
            vTo[0] = vv
            pin[0] = F(0) if first_pass else ld(p2, q, row)
            fl[0] = ld(f2, q, row) & 15

        prep_k0(0)
        for t in range(T):
            f_eff = fl.copy()
            f_eff[k_act:] = 0                            # inactive iterations pass through
            uCf = np.empty((KB, LANES), F); uRp = np.empty((KB, LANES), F)
            vCf = np.empty((KB, LANES), F); vTp = np.empty((KB, LANES), F)
            pout = np.empty((KB, LANES), F)
            for k in range(KB):
                uCf[k], uRp[k], vCf[k], vTp[k], pout[k] = cell_update_vec(
                    f_eff[k], omega, pc, uL[k], uRo[k], vB[k], vTo[k], pin[k])
            # final values of the pass leave the pipeline after iteration KB-1
            col, row = q - (KB - 1), t - L - (KB - 1) - 1
            ok = (col >= 0) & (col < W) & (row >= 0) & (row < H)
            uo2[col[ok], row[ok]] = uCf[KB - 1][ok]
            vo2[col[ok], row[ok]] = vCf[KB - 1][ok]
            po2[col[ok], row[ok]] = pout[KB - 1][ok]
            # record of the last lane for band b+1 (and, in CUDA, of lane 31 for the next warp)
            rec.append((uRp[:, -1].copy(), vCf[:, -1].copy(), pout[:, -1].copy(), fl[:, -1].copy()))
            # ---- shift: operands of step t+1 ------------------------------------------------
            # Each band runs on its own clock; band b's lanes continue the diagonal of band b-1,
            # so local step t here is local step t + LANES there: the record consumed now is
            # the one band b-1 wrote at ITS step t + LANES (band b trails by >= LANES steps).
            if prev_rec is not None and t + LANES < len(prev_rec):
                l_uRp, l_vCf, l_pout, l_fl = prev_rec[t + LANES]
            else:
                l_uRp = l_vCf = l_pout = np.zeros(KB, F)
                l_fl = np.zeros(KB, np.uint8)

            def shfl_up(x, left):                        # x: (KB, LANES); left: (KB,) for lane 0
                y = np.empty_like(x)
                y[:, 1:] = x[:, :-1]
                y[:, 0] = left
                return y

            n_uL = shfl_up(uRp, l_uRp)
            s_vCf = shfl_up(vCf, l_vCf)
            s_pout = shfl_up(pout, l_pout)
            s_fl = shfl_up(fl, l_fl)
            vB[:] = vTp
            uRo[1:] = uCf[:-1]
            uL[:] = n_uL
            vTo[1:] = s_vCf[:-1]
            pin[1:] = pd[1:]                             # two-step path: pd was filled last step
            fl[1:] = fld[1:]
            pd[1:] = s_pout[:-1]
            fld[1:] = s_fl[:-1]
            prep_k0(t + 1)
        prev_rec = rec
    return uo, vo, po


def systolic_project(u, v, flags, W, H, iterations, KB, LANES, omega, pc, gdt):
    """make_incompressible (+ fused add_gravity) as the CUDA kernel schedules it."""
    p = np.zeros(W * H, F)
    done, first = 0, True
    while done < iterations or first:
        k_act = min(KB, iterations - done)
        u, v, p = systolic_pass(u, v, p, flags, W, H, KB, k_act, LANES, omega, pc, gdt, first)
        done += k_act
        first = False
        if iterations == 0:
            break
    return u, v, p
