"""CPU oracle for the MGVI hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy float64 restatement of the algorithm of bat/MGVI.jl v0.4.5
(`/root/reference/src/*.jl`) for the path SURVEY.md section 8 names.  It is the checker the
CUDA library is compared with; it is never the thing that is shipped or measured:

* only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs
  of `bench.py` may import it;
* nothing under `mgvi.jl_b200/` imports it, and the product path raises when the CUDA
  library is missing instead of falling back to this code.

Pinning status: the reference's own tests hold known answers only for
`fisher_information` (test/information/test_information.jl:26-28,53-74), the Jacobian
strategy consistency (test/test_jacobians.jl:28-32,45-49) and the statistical
dense-vs-CG sampler equivalence (test/test_samplers.jl:27-38).  The oracle is checked
against all of those in `tests/test_oracle.py`.  The reference cannot be executed here
(no Julia in the image, none on the GPU box), and no reference test pins any output of
`mgvi_step`, the sampler CG, NewtonCG or the line search, so for those rows the header
says it plainly: **parity unpinned** -- it rests on this restatement of the published
algorithms of Krylov.jl `cg!`, IterativeSolvers `cg_iterator!` and LineSearches
`StrongWolfe` (un-vendored dependencies, bounded by Project.toml:40-67, not pinned).
"""
from .mgvi_oracle import *  # noqa: F401,F403
