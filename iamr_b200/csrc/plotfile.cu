// plotfile.cu -- AMReX plotfile (VisMF) writer for the level state: the on-disk format IAMR produces through
// Amr::writePlotFile -> AmrLevel::writePlotFile (NavierStokesBase.cpp:3343-3352 thePlotFileType "NavierStokes-V1.1";
// NavierStokes.cpp:1080-1196 writePlotFilePre/Post; layout described in Docs/sphinx_documentation/source/Software.rst:712-750:
// one folder per plotfile, one sub-folder per level, each MultiFab = a header file plus data files the ranks append their FABs
// to).  Host I/O code: it lets a site with a real IAMR/AMReX build compare this library's results with `fcompare` / yt /
// Amrvis -- the route to pinning parity that this repository cannot take itself.
//
// Files written for `dir`:
//   dir/Header                  plotfile header (variable names, geometry, per-level grids)
//   dir/Level_0/Cell_H          VisMF header (version 1: box array, FabOnDisk offsets, per-FAB min / max)
//   dir/Level_0/Cell_D_<rank>   the FABs owned by each rank, in box order: "FAB <RealDescriptor><box> <ncomp>\n" + raw doubles
//   dir/job_info                the free-form text file of writePlotFilePost
// Plot variables = the cell-centred state components AmrLevel::setPlotVariables registers (NS_setup.cpp:253-269,339-360):
// x_velocity y_velocity z_velocity density tracer gradpx gradpy gradpz.
#include <sys/stat.h>
#include <cerrno>
#include <fstream>
#include <sstream>
#include "solvers.h"

namespace ix {

namespace {

#define IX_TRY(call) do { int rc_ = (call); if (rc_ != IAMRX_OK) return rc_; } while (0)

std::string box_str(const Bx& b) {
  std::ostringstream o;
  o << "((" << b.lo[0] << "," << b.lo[1] << "," << b.lo[2] << ") (" << b.hi[0] << "," << b.hi[1] << "," << b.hi[2] << ") (0,0,0))";
  return o.str();
}
// amrex::RealDescriptor of native IEEE little-endian doubles, as FArrayBox::writeOn prints it
const char* REAL_DESC = "((8, (64 11 52 0 1 12 0 1023)),(8, (8 7 6 5 4 3 2 1)))";
std::string fab_header(const Bx& b, int ncomp) {
  std::ostringstream o;
  o << "FAB " << REAL_DESC << box_str(b) << " " << ncomp << "\n";
  return o.str();
}
std::string rank_file(int rank) {
  char buf[32];
  snprintf(buf, sizeof(buf), "Cell_D_%05d", rank);
  return buf;
}
int make_dir(const std::string& d) {
  if (mkdir(d.c_str(), 0755) != 0 && errno != EEXIST) { set_error("plotfile: cannot create directory " + d); return IAMRX_ERR_ARG; }
  return IAMRX_OK;
}

}  // namespace

// fields: list of (MF, first comp, ncomp) concatenated into the plot MultiFab; names: one per component
int write_plotfile(Level& L, const std::vector<const MF*>& mfs, const std::vector<int>& comp0, const std::vector<int>& ncomps,
                   const std::vector<std::string>& names, const char* dir, const char* plot_type, double time, int level_steps,
                   cudaStream_t s) {
  const int ncomp = (int)names.size();
  const int me = comm().rank, nranks = comm().nranks;
  const std::string root(dir), lev_dir = root + "/Level_0";
  IX_TRY(make_dir(root));
  IX_TRY(make_dir(lev_dir));
  const int nb = (int)L.boxes.size();
  // every rank can compute every FAB's size, hence the offsets inside each rank's data file, without communication
  std::vector<long long> offset(nb, 0), next(nranks, 0);
  for (int b = 0; b < nb; ++b) {
    const int r = L.owner[b];
    offset[b] = next[r];
    next[r] += (long long)fab_header(L.boxes[b], ncomp).size() + (long long)L.boxes[b].npts() * ncomp * (long long)sizeof(double);
  }
  // per-FAB min / max of every component (VisMF header): local boxes on the device, then a min / max all-reduce
  std::vector<double> mn((size_t)nb * ncomp, 1.0e300), mx((size_t)nb * ncomp, -1.0e300);
  {
    std::ofstream df(lev_dir + "/" + rank_file(me), std::ios::binary | std::ios::trunc);
    if (L.nlocal() > 0 && !df) { set_error("plotfile: cannot open data file in " + lev_dir); return IAMRX_ERR_ARG; }
    std::vector<double> host;
    for (int il = 0; il < L.nlocal(); ++il) {
      const int b = L.local[il];
      const Bx& bx = L.boxes[b];
      const size_t npts = (size_t)bx.npts();
      double* dbuf = dev_alloc(npts * ncomp);
      if (!dbuf) return IAMRX_ERR_CUDA;
      int c = 0;
      for (size_t m = 0; m < mfs.size(); ++m) {
        const int rc = k::pack(bx, dbuf + (size_t)c * npts, mfs[m]->c(il, comp0[m]), ncomps[m], s);
        if (rc) { dev_free(dbuf); return rc; }
        c += ncomps[m];
      }
      host.resize(npts * ncomp);
      cudaError_t e = cudaMemcpyAsync(host.data(), dbuf, npts * ncomp * sizeof(double), cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      dev_free(dbuf);
      if (e != cudaSuccess) { set_error(std::string("plotfile: ") + cudaGetErrorString(e)); return IAMRX_ERR_CUDA; }
      for (int n = 0; n < ncomp; ++n) {
        double lo = 1.0e300, hi = -1.0e300;
        const double* p = host.data() + (size_t)n * npts;
        for (size_t q = 0; q < npts; ++q) { lo = p[q] < lo ? p[q] : lo; hi = p[q] > hi ? p[q] : hi; }
        mn[(size_t)b * ncomp + n] = lo; mx[(size_t)b * ncomp + n] = hi;
      }
      const std::string h = fab_header(bx, ncomp);
      df.write(h.data(), (std::streamsize)h.size());
      df.write(reinterpret_cast<const char*>(host.data()), (std::streamsize)(npts * ncomp * sizeof(double)));
    }
    if (L.nlocal() > 0 && !df) { set_error("plotfile: write failed"); return IAMRX_ERR_ARG; }
  }
  if (nranks > 1) {
    const size_t n = (size_t)nb * ncomp;
    double* d = dev_alloc(2 * n);
    if (!d) return IAMRX_ERR_CUDA;
    std::vector<double> both(2 * n);
    for (size_t q = 0; q < n; ++q) { both[q] = mn[q]; both[n + q] = -mx[q]; }   // one min-reduction for both
    cudaMemcpyAsync(d, both.data(), 2 * n * sizeof(double), cudaMemcpyHostToDevice, s);
    int rc = IAMRX_OK;
    for (size_t q0 = 0; q0 < 2 * n && rc == IAMRX_OK; q0 += 1 << 20) rc = comm_allreduce(d + q0, (int)std::min<size_t>(1 << 20, 2 * n - q0), 1, s);
    cudaMemcpyAsync(both.data(), d, 2 * n * sizeof(double), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    dev_free(d);
    if (rc) return rc;
    for (size_t q = 0; q < n; ++q) { mn[q] = both[q]; mx[q] = -both[n + q]; }
  }
  if (me != 0) return IAMRX_OK;
  // ---- Level_0/Cell_H ------------------------------------------------------------------------------------------------
  {
    std::ofstream h(lev_dir + "/Cell_H", std::ios::trunc);
    h.precision(17);
    h << 1 << "\n" << 0 << "\n" << ncomp << "\n" << 0 << "\n";
    h << "(" << nb << " 0\n";
    for (int b = 0; b < nb; ++b) h << box_str(L.boxes[b]) << "\n";
    h << ")\n";
    h << nb << "\n";
    for (int b = 0; b < nb; ++b) h << "FabOnDisk: " << rank_file(L.owner[b]) << " " << offset[b] << "\n";
    h << "\n";
    for (int pass = 0; pass < 2; ++pass) {
      const std::vector<double>& v = pass == 0 ? mn : mx;
      h << nb << "," << ncomp << "\n";
      for (int b = 0; b < nb; ++b) {
        for (int n = 0; n < ncomp; ++n) h << v[(size_t)b * ncomp + n] << ",";
        h << "\n";
      }
      h << "\n";
    }
    if (!h) { set_error("plotfile: cannot write Cell_H"); return IAMRX_ERR_ARG; }
  }
  // ---- Header -----------------------------------------------------------------------------------------------------------
  {
    std::ofstream h(root + "/Header", std::ios::trunc);
    h.precision(17);
    const iamrx_geom& g = L.geom;
    h << plot_type << "\n" << ncomp << "\n";
    for (const std::string& nm : names) h << nm << "\n";
    h << 3 << "\n" << time << "\n" << 0 << "\n";
    double phi[3];
    for (int d = 0; d < 3; ++d) phi[d] = g.prob_lo[d] + (g.domain.hi[d] - g.domain.lo[d] + 1) * g.dx[d];
    h << g.prob_lo[0] << " " << g.prob_lo[1] << " " << g.prob_lo[2] << " \n";
    h << phi[0] << " " << phi[1] << " " << phi[2] << " \n";
    h << "\n";                                       // refinement ratios: none for a single level
    h << box_str(L.domain) << " \n";
    h << level_steps << " \n";
    h << g.dx[0] << " " << g.dx[1] << " " << g.dx[2] << " \n";
    h << 0 << "\n" << 0 << "\n";                       // coordinate system (Cartesian), boundary width
    h << 0 << " " << nb << " " << time << "\n" << level_steps << "\n";
    for (int b = 0; b < nb; ++b)
      for (int d = 0; d < 3; ++d)
        h << g.prob_lo[d] + (L.boxes[b].lo[d] - g.domain.lo[d]) * g.dx[d] << " " << g.prob_lo[d] + (L.boxes[b].hi[d] + 1 - g.domain.lo[d]) * g.dx[d] << "\n";
    h << "Level_0/Cell\n";
    if (!h) { set_error("plotfile: cannot write Header"); return IAMRX_ERR_ARG; }
  }
  {
    std::ofstream j(root + "/job_info", std::ios::trunc);
    j << "===============================================================================\n Job Information\n"
      << "===============================================================================\n"
      << "number of MPI processes: " << nranks << "\n\n"
      << "written by libiamrx (iamr-b200), plotfile type " << plot_type << "\n";
  }
  return IAMRX_OK;
}

}  // namespace ix
