// dump_reference_vectors.cc -- pins the oracle against the REAL reference (dealii-X/portable-multigrid).
//
// This repository's oracle (oracle/*.c) is a restatement of the reference's algorithm; deal.II >= 9.8, Kokkos and MPI are not
// installed where it was written, so its parity with the reference is pinned only indirectly (tests/golden/README.md).  Anyone
// with a deal.II/Kokkos build can close that gap: compile this program against the reference's headers, run it, and feed its
// output to tests/golden/dealii_dump/compare_with_reference_dump.py.
//
// What it does, for a uniformly refined unit cube with homogeneous Dirichlet values on boundary id 0 (the reference drivers'
// mesh, source/geometric_multigrid/program.cc:130,163-166):
//   * builds the reference's level operators / transfers / Chebyshev smoothers / V-cycle exactly as its driver does
//     (program.cc:205-285), with the driver's parameters;
//   * numbers every DoF LEXICOGRAPHICALLY by its support point (x fastest) -- the numbering of this repository -- so that
//     vectors can be compared entry by entry whatever numbering deal.II chose;
//   * applies LaplaceOperator::vmult, GeometricTransfer::prolongate_and_add / restrict_and_add and VCycleMultigrid::vmult to
//     the synthetic vector tests/helpers.py:splitmix_src defines (2 u01(splitmix64(i ^ GOLDEN)) - 1, i = lexicographic index),
//     and runs the driver's CG solve (program.cc:342-355);
//   * writes  <prefix>_{src,vmult,prolongated,restricted,vcycle}.f64  (raw little-endian doubles, lexicographic order) and
//     <prefix>_cg.txt (iteration count, then one residual norm per line).
//
// Build (sketch): add_executable(dump dump_reference_vectors.cc); target_include_directories(dump PRIVATE <reference>/include);
// deal_ii_setup_target(dump)  -- with the same deal.II / Kokkos configuration as the reference's own drivers.
// Run: ./dump <refinements> <prefix>     (fe_degree is the template parameter DEGREE below, default 2).
#include <deal.II/base/quadrature_lib.h>
#include <deal.II/distributed/tria.h>
#include <deal.II/dofs/dof_tools.h>
#include <deal.II/fe/fe_q.h>
#include <deal.II/fe/mapping_q1.h>
#include <deal.II/grid/grid_generator.h>
#include <deal.II/lac/la_parallel_vector.h>
#include <deal.II/lac/precondition.h>
#include <deal.II/lac/solver_cg.h>
#include <deal.II/multigrid/mg_transfer_global_coarsening.h>
#include <deal.II/numerics/vector_tools.h>

#include <base/portable_laplace_operator_base.h>
#include <base/portable_mg_transfer_base.h>
#include <multigrid/portable_geometric_transfer.h>
#include <multigrid/portable_v_cycle_multigrid.h>
#include <operators/portable_laplace_operator.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <map>

using namespace dealii;
#ifndef DEGREE
#define DEGREE 2
#endif
constexpr int dim = 3;
using DeviceVector = LinearAlgebra::distributed::Vector<double, MemorySpace::Default>;
using HostVector   = LinearAlgebra::distributed::Vector<double, MemorySpace::Host>;

// tests/helpers.py:splitmix_src
static double splitmix(std::uint64_t i, std::uint64_t salt)
{
  std::uint64_t z = ((i + salt * 0x632BE59BD9B4E019ull) ^ 0x9E3779B97F4A7C15ull) + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return 2.0 * (double)(z >> 11) * (1.0 / 9007199254740992.0) - 1.0;
}

// lexicographic index of every DoF from its support point: rank of its x, y, z among the distinct coordinates
static std::vector<std::uint64_t> lexicographic_numbering(const DoFHandler<dim> &dh)
{
  const auto pts = DoFTools::map_dofs_to_support_points(MappingQ1<dim>(), dh);
  std::array<std::vector<double>, dim> axis;
  for (const auto &kv : pts)
    for (int d = 0; d < dim; ++d) axis[d].push_back(kv.second[d]);
  for (auto &a : axis)
    {
      std::sort(a.begin(), a.end());
      a.erase(std::unique(a.begin(), a.end(), [](double x, double y) { return std::abs(x - y) < 1e-10; }), a.end());
    }
  std::vector<std::uint64_t> lex(dh.n_dofs());
  for (const auto &kv : pts)
    {
      std::uint64_t idx = 0, stride = 1;
      for (int d = 0; d < dim; ++d)
        {
          const auto it = std::lower_bound(axis[d].begin(), axis[d].end(), kv.second[d] - 1e-10);
          idx += stride * (std::uint64_t)(it - axis[d].begin());
          stride *= axis[d].size();
        }
      lex[kv.first] = idx;
    }
  return lex;
}

static void to_device(DeviceVector &dst, const std::vector<double> &lexvals, const std::vector<std::uint64_t> &lex)
{
  HostVector h(dst.get_partitioner());
  for (const auto i : h.locally_owned_elements()) h[i] = lexvals[lex[i]];
  LinearAlgebra::ReadWriteVector<double> rw(h.locally_owned_elements());
  rw.import_elements(h, VectorOperation::insert);
  dst.import_elements(rw, VectorOperation::insert);
}

static void dump(const std::string &name, const DeviceVector &v, const std::vector<std::uint64_t> &lex)
{
  LinearAlgebra::ReadWriteVector<double> rw(v.locally_owned_elements());
  rw.import_elements(v, VectorOperation::insert);
  std::vector<double> out(lex.size(), 0.0);
  for (const auto i : v.locally_owned_elements()) out[lex[i]] = rw[i];
  std::ofstream f(name, std::ios::binary);
  f.write(reinterpret_cast<const char *>(out.data()), out.size() * sizeof(double));
}

int main(int argc, char **argv)
{
  Utilities::MPI::MPI_InitFinalize mpi(argc, argv, 1); // run on ONE rank: the dump gathers nothing
  const unsigned int refinements = argc > 1 ? std::atoi(argv[1]) : 3;
  const std::string  prefix      = argc > 2 ? argv[2] : "ref";
  parallel::distributed::Triangulation<dim> tria(MPI_COMM_WORLD);
  GridGenerator::hyper_cube(tria, 0., 1.);
  tria.refine_global(refinements);
  const auto coarse = MGTransferGlobalCoarseningTools::create_geometric_coarsening_sequence(tria);
  const FE_Q<dim> fe(DEGREE);
  const unsigned int L = coarse.size() - 1;
  MGLevelObject<DoFHandler<dim>>           dhs(0, L);
  MGLevelObject<AffineConstraints<double>> cons(0, L);
  Functions::ZeroFunction<dim> zero;
  const std::map<types::boundary_id, const Function<dim> *> bc = {{0, &zero}};
  for (unsigned int l = 0; l <= L; ++l)
    {
      dhs[l].reinit(*coarse[l]);
      dhs[l].distribute_dofs(fe);
      cons[l].reinit(dhs[l].locally_owned_dofs(), DoFTools::extract_locally_relevant_dofs(dhs[l]));
      VectorTools::interpolate_boundary_values(dhs[l], bc, cons[l]);
      cons[l].close();
    }
  MGLevelObject<std::unique_ptr<Portable::LaplaceOperatorBase<dim, double>>> A(0, L);
  for (unsigned int l = 0; l <= L; ++l)
    A[l] = std::make_unique<Portable::LaplaceOperator<dim, DEGREE, double>>(dhs[l], cons[l], false);
  MGLevelObject<std::unique_ptr<Portable::MGTransferBase<dim, double>>> T(0, L);
  for (unsigned int l = 1; l <= L; ++l)
    {
      T[l] = std::make_unique<Portable::GeometricTransfer<dim, DEGREE, double>>();
      T[l]->reinit(A[l - 1]->get_matrix_free(), A[l]->get_matrix_free(), cons[l - 1], cons[l]);
    }
  using Smoother = PreconditionChebyshev<Portable::LaplaceOperatorBase<dim, double>, DeviceVector>;
  MGLevelObject<Smoother> S(0, L);
  for (unsigned int l = 0; l <= L; ++l)
    {
      typename Smoother::AdditionalData d;
      d.smoothing_range     = l > 0 ? 15. : 1e-3;
      d.degree              = l > 0 ? 5 : numbers::invalid_unsigned_int;
      d.eig_cg_n_iterations = l > 0 ? 10 : A[0]->m();
      A[l]->compute_diagonal();
      d.preconditioner = A[l]->get_matrix_diagonal_inverse();
      S[l].initialize(*A[l], d);
    }
  const auto lexf = lexicographic_numbering(dhs[L]);
  const auto lexc = lexicographic_numbering(dhs[L - 1]);
  std::vector<double> src(lexf.size()), srcc(lexc.size());
  for (std::uint64_t i = 0; i < src.size(); ++i) src[i] = splitmix(i, 0);
  for (std::uint64_t i = 0; i < srcc.size(); ++i) srcc[i] = splitmix(i, 5);
  DeviceVector u, Au, uc, rc;
  A[L]->initialize_dof_vector(u);
  A[L]->initialize_dof_vector(Au);
  A[L - 1]->initialize_dof_vector(uc);
  A[L - 1]->initialize_dof_vector(rc);
  to_device(u, src, lexf);
  dump(prefix + "_src.f64", u, lexf);
  A[L]->vmult(Au, u);
  dump(prefix + "_vmult.f64", Au, lexf);
  to_device(uc, srcc, lexc);
  Au = 0.;
  T[L]->prolongate_and_add(Au, uc);
  dump(prefix + "_prolongated.f64", Au, lexf);
  rc = 0.;
  T[L]->restrict_and_add(rc, u);
  dump(prefix + "_restricted.f64", rc, lexc);
  Portable::VCycleMultigrid<dim, double, Portable::MGTransferBase<dim, double>> mg(A, T, S, 2, 2);
  // the V-cycle acts on residuals: zero on constrained DoFs
  std::vector<double> res = src;
  for (const auto i : u.locally_owned_elements())
    if (cons[L].is_constrained(i)) res[lexf[i]] = 0.;
  to_device(u, res, lexf);
  mg.vmult(Au, u);
  dump(prefix + "_vcycle.f64", Au, lexf);
  // the driver's solve: right-hand side f = 1, CG to 1e-12 |b| preconditioned by one V-cycle (program.cc:289-355)
  DeviceVector b, x;
  A[L]->initialize_dof_vector(b);
  A[L]->initialize_dof_vector(x);
  {
    HostVector bh(b.get_partitioner());
    VectorTools::create_right_hand_side(dhs[L], QGauss<dim>(DEGREE + 1), Functions::ConstantFunction<dim>(1.), bh, cons[L]);
    LinearAlgebra::ReadWriteVector<double> rw(bh.locally_owned_elements());
    rw.import_elements(bh, VectorOperation::insert);
    b.import_elements(rw, VectorOperation::insert);
  }
  SolverControl control(b.size(), 1e-12 * b.l2_norm());
  control.enable_history_data();
  SolverCG<DeviceVector> cg(control);
  cg.solve(*A[L], x, b, mg);
  std::ofstream f(prefix + "_cg.txt");
  f.precision(17);
  f << control.last_step() << "\n";
  for (const double r : control.get_history_data()) f << r << "\n";
  return 0;
}
