"""``SolverOptions``: same fields, defaults and validation errors as the reference's
dataclass (tdgl/solver/options.py:19-166), plus B200-only knobs appended at the end with
defaults so that existing scripts run unchanged.

Differences (single backend): ``gpu`` and ``sparse_solver`` are accepted for
compatibility but do not select anything — the step always runs on the CUDA engine and
the mu system is always solved by its on-device AMG-preconditioned CG.
"""

from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Optional, Union


class SolverOptionsError(ValueError):
    pass


class SparseSolver(Enum):
    """Sparse solvers the reference supports (options.py:10-16); accepted, ignored."""

    SUPERLU = "superlu"
    UMFPACK = "umfpack"
    PARDISO = "pardiso"
    CUPY = "cupy"


@dataclass
class SolverOptions:
    solve_time: float
    skip_time: float = 0.0
    dt_init: float = 1e-6
    dt_max: float = 1e-1
    adaptive: bool = True
    adaptive_window: int = 10
    max_solve_retries: int = 10
    adaptive_time_step_multiplier: float = 0.25
    output_file: Union[str, None] = None
    terminal_psi: Union[float, complex, None] = 0.0
    gpu: bool = False
    sparse_solver: Union[SparseSolver, str] = SparseSolver.SUPERLU
    pause_on_interrupt: bool = True
    save_every: int = 100
    progress_interval: int = 0
    monitor: bool = False
    monitor_update_interval: float = 1.0
    field_units: str = "mT"
    current_units: str = "uA"
    include_screening: bool = False
    max_iterations_per_step: int = 1000
    screening_tolerance: float = 1e-3
    screening_step_size: float = 0.1
    screening_step_drag: float = 0.5
    # ---- B200 engine knobs (no counterpart in the reference) ----------------------------
    mu_rtol: float = 1e-10        # relative residual of the on-device mu solve
    mu_max_iterations: int = 500
    cuda_device: Optional[int] = None   # None: device 0, or (distributed) the rank's
    #                               current CUDA device (torch.cuda.current_device())
    use_cuda_graph: bool = True   # device-side step / retry / CG loops in one CUDA graph
    async_save: bool = True       # saves go through pinned double buffers + a writer thread
    distributed: bool = False     # True: this process is one shard of a torchrun job (the
    #                               mesh is domain-decomposed over torch.distributed's ranks)

    def validate(self) -> None:
        """Same checks and messages as the reference (options.py:91-166)."""
        if self.dt_init > self.dt_max:
            raise SolverOptionsError("dt_init must be less than or equal to dt_max.")
        if self.terminal_psi is not None and not (0 <= abs(self.terminal_psi) <= 1):
            raise SolverOptionsError(
                "terminal_psi must be None or have absolute value in [0, 1]"
                f" (got {self.terminal_psi}).")
        if not (0 < self.adaptive_time_step_multiplier < 1):
            raise SolverOptionsError(
                "adaptive_time_step_multiplier must be in (0, 1)"
                f" (got {self.adaptive_time_step_multiplier}).")
        if not (0 < self.screening_step_drag <= 1):
            raise SolverOptionsError(
                f"screening_step_drag must be in (0, 1] (got {self.screening_step_drag}).")
        if self.screening_step_size <= 0:
            raise SolverOptionsError(
                f"screening_step_size must be in > 0 (got {self.screening_step_size}).")
        if self.screening_tolerance <= 0:
            raise SolverOptionsError(
                f"screening_tolerance must be in > 0 (got {self.screening_tolerance}).")
        solver = self.sparse_solver
        if isinstance(solver, str):
            try:
                solver = SparseSolver[solver.upper()]
            except KeyError:
                valid = list(SparseSolver.__members__.keys())
                raise SolverOptionsError(
                    f"sparse solver must be one of {valid!r}, got {solver}.")
            self.sparse_solver = solver
        if self.include_screening and self.distributed:
            raise SolverOptionsError(
                "include_screening=True is not available with distributed=True: the induced"
                " vector potential is an all-pairs sum over the whole mesh.")
        if not (self.mu_rtol > 0):
            raise SolverOptionsError(f"mu_rtol must be > 0 (got {self.mu_rtol}).")
        if self.adaptive_window < 1 or self.adaptive_window > 1024:
            raise SolverOptionsError("adaptive_window must be in [1, 1024].")
