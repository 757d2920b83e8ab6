"""learnmultigrid_b200 -- B200-native (sm_100a) multigrid V-cycle engine behind the LearnMultigrid API.

Drop-in module layout (same sub-module and class names as the reference's `learn_multigrid` package):
    learnmultigrid_b200.solvers.{Solver,Jacobi,GaussSeidel,CG,Multigrid}
    learnmultigrid_b200.L2_projection.{L2Projection,CouplingOperator,Intersection}
    learnmultigrid_b200.assembly.{MassMatrix,StiffnessMatrix,LoadVector,LoadFunction,Quadrature,ShapeFunction,
                                  MapReferenceElement}
    learnmultigrid_b200.mesh.{Mesh1D,Mesh2D}
The top-level `learn_multigrid` package in this repository re-exports them under the reference's import paths.

All solve-phase compute runs in hand-written CUDA kernels (learnmultigrid_b200/csrc, C ABI in include/mgb200.h);
there is no CPU fallback.
"""
__version__ = "0.1.0"
