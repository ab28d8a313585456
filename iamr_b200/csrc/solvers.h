// solvers.h -- level solvers built from the multigrid drivers: what
// Hydro::MacProjector, Hydro::NodalProjector and the MLABecLaplacian /
// MLTensorOp + MLMG pairs are to IAMR (MacProj.cpp:1084-1184,
// Projection.cpp:2385-2567, Diffusion.cpp:207-957).  Multigrid hierarchies are
// cached per level (the reference rebuilds LPInfo/MLMG objects on every call;
// SURVEY.md section 8b "Ownership").
#pragma once
#include <map>
#include "mlmg.h"

namespace ix {

struct LevelSolvers {
  std::unique_ptr<CellMG> mac;
  int mac_mc = -1;
  MF mac_beta[3];
  MF mac_rhs;
  std::map<int, std::unique_ptr<CellMG>> diff;  // key = ncomp*2 + tensor
  std::map<int, int> diff_mc;
  std::unique_ptr<NodeMG> nodal;
  int nodal_mc = -1;
  MF nodal_rhs;
};

int mac_project(Level& L, LevelSolvers& sv, MF U[3], const MF& rho, const MF* rhs, MF& phi,
                double rhs_scale, iamrx_mg_info* info, cudaStream_t s, const k::LinBC* bc = nullptr);
int mac_get_fluxes(Level& L, LevelSolvers& sv, MF F[3], MF& phi, cudaStream_t s);
// keep_dirichlet: the nodes ON Dirichlet domain sides keep the values of phi on entry (Projection::set_outflow_bcs has put the
// hydrostatic pressure there) instead of being held at zero
int nodal_project(Level& L, LevelSolvers& sv, MF& vel, const MF& sigma, MF& phi, MF* gp,
                  int increment_gp, iamrx_mg_info* info, cudaStream_t s, const k::NodalBC* bc = nullptr, bool keep_dirichlet = false);
int diffusion_apply(Level& L, LevelSolvers& sv, bool tensor, int ncomp, MF& out, MF& soln, double a,
                    double b, const MF* acoef, MF eta[3], cudaStream_t s, const k::LinBC* bc = nullptr);
int diffusion_solve(Level& L, LevelSolvers& sv, bool tensor, int ncomp, MF& soln, const MF& rhs,
                    double a, double b, const MF* acoef, MF eta[3], iamrx_mg_info* info, cudaStream_t s,
                    const k::LinBC* bc = nullptr);
k::LinBC periodic_linbc();
// MLNodeLaplacian::compGrad (NSB.cpp:4106-4118): gp = grad(p) on cell centres
int comp_grad(Level& L, MF& gp, MF& p, cudaStream_t s);

}  // namespace ix
