// qkxtm_invert_test.cpp -- a small driver in the mould of qkxtm/MG_Bench.cpp / qkxtm/Calc_Loops.cpp main():
// parse the hot-path flags (the ~15 of SURVEY.md section 5 / Appendix D, same names), fill QudaGaugeParam /
// QudaInvertParam / qudaQKXTMinfo the way the reference drivers do (qkxtm/Calc_Loops.cpp:189-225,380-497),
// build a synthetic gauge field, then initQuda -> init_qudaQKXTM -> loadGaugeQuda -> one entry point.
// Results go to a raw binary file so that the parity tests can check them against the CPU oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>
#include "../../../include/qudaQKXTM_tmq.h"
#include "../../../include/tmq_host.h"

using namespace quda;

static void usage() {
  printf("qkxtm_invert_test --dim X Y Z T [--kappa k | --mass m] [--mu mu] [--tol t] [--niter n]\n"
         "   [--prec-sloppy double|single] [--recon 12|18] [--matpc even-even|odd-odd] [--mass-normalization kappa|mass]\n"
         "   [--test invert|mgbench|loops|mdagm|mat] [--source z4|gaussian] [--seed s] [--verbosity-level silent|summarize|verbose]\n"
         "   [--out file] [--dump-inputs prefix] [--gridsize gx gy gz gt]  (--dim is the LOCAL lattice; one process per rank)\n"
         "   --test twop [--Q_sq q] [--src x y z t] [--nsmearGauss n --alphaGauss a]: meson two-point function, written to\n"
         "       <out>.mesons.SS.xx.yy.zz.tt.dat; with --tsink dt [--proj 0..4] [--particle proton|neutron] also the three-point function\n");
}

int main(int argc, char **argv) {
  int dim[4] = {8, 8, 8, 16};
  double kappa = -1.0, mass = 0.1, mu = 0.1, tol = 1e-7;      // defaults: qkxtm/QKXTM_util.cpp:1594-1600
  int niter = 5000, recon = 18;
  std::string prec_sloppy = "double", matpc = "even-even", test = "invert", source = "z4", out, massnorm = "kappa", verb = "summarize";
  unsigned long long seed = 100;
  int nev = 4, nkv = 16, polydeg = 20;
  double amin = 0.385, amax = 2.0, eig_tol = 1e-10, csw = 0.0;
  int q_sq = 0, src_pos[4] = {0, 0, 0, 0}, tsink = 0, proj = 0;
  int grid[4] = {1, 1, 1, 1};
  std::string particle = "proton";
  std::string dslash_type = "twisted-mass";
  int Nstoch = 1, NdumpStep = 1, k_probing = 0, hadamLow = 0, hadamHigh = 0, n_defl_steps = 0, defl_step_nEv[MAX_DEFLSTEPS] = {0};
  bool spinColorDil = false, isFullOp = false;
  std::string source_type = "random";
  int nsrc = 12, e2e_reps = 2;
  int nsmearGauss = 0; double alphaGauss = 4.0;                              // qkxtm/QKXTM_util.cpp:1652-1654
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto need = [&](int n) { if (i + n >= argc) { usage(); exit(2); } };
    if (a == "--dim") { need(4); for (int d = 0; d < 4; d++) dim[d] = atoi(argv[++i]); }
    else if (a == "--kappa") { need(1); kappa = atof(argv[++i]); }
    else if (a == "--mass") { need(1); mass = atof(argv[++i]); }
    else if (a == "--mu") { need(1); mu = atof(argv[++i]); }
    else if (a == "--tol") { need(1); tol = atof(argv[++i]); }
    else if (a == "--niter") { need(1); niter = atoi(argv[++i]); }
    else if (a == "--prec-sloppy") { need(1); prec_sloppy = argv[++i]; }
    else if (a == "--recon") { need(1); recon = atoi(argv[++i]); }
    else if (a == "--matpc") { need(1); matpc = argv[++i]; }
    else if (a == "--mass-normalization") { need(1); massnorm = argv[++i]; }
    else if (a == "--test") { need(1); test = argv[++i]; }
    else if (a == "--source") { need(1); source = argv[++i]; }
    else if (a == "--seed") { need(1); seed = strtoull(argv[++i], nullptr, 10); }
    else if (a == "--verbosity-level") { need(1); verb = argv[++i]; }
    else if (a == "--out") { need(1); out = argv[++i]; }
    else if (a == "--dslash-type") { need(1); dslash_type = argv[++i]; }      // twisted-mass | twisted-clover (qkxtm/misc.cpp:830-859)
    else if (a == "--nsmearGauss") { need(1); nsmearGauss = atoi(argv[++i]); }
    else if (a == "--alphaGauss") { need(1); alphaGauss = atof(argv[++i]); }
    else if (a == "--csw") { need(1); csw = atof(argv[++i]); }                 // qkxtm/QKXTM_util.cpp:1642
    else if (a == "--PolyDeg") { need(1); polydeg = atoi(argv[++i]); }        // the reference's ARPACK flags (qkxtm/QKXTM_util.cpp)
    else if (a == "--nEv") { need(1); nev = atoi(argv[++i]); }
    else if (a == "--nKv") { need(1); nkv = atoi(argv[++i]); }
    else if (a == "--amin") { need(1); amin = atof(argv[++i]); }
    else if (a == "--amax") { need(1); amax = atof(argv[++i]); }
    else if (a == "--tolArpack") { need(1); eig_tol = atof(argv[++i]); }
    else if (a == "--Q_sq") { need(1); q_sq = atoi(argv[++i]); }               // qkxtm/QKXTM_util.cpp (momenta with p^2 <= Q_sq)
    else if (a == "--src") { need(4); for (int d = 0; d < 4; d++) src_pos[d] = atoi(argv[++i]); }
    else if (a == "--gridsize") { need(4); for (int d = 0; d < 4; d++) grid[d] = atoi(argv[++i]); }   // process grid x y z t (qkxtm/QKXTM_util.cpp:2198); --dim is LOCAL
    else if (a == "--tsink") { need(1); tsink = atoi(argv[++i]); }             // > 0: also the three-point function at this sink-source separation
    else if (a == "--proj") { need(1); proj = atoi(argv[++i]); }               // WHICHPROJECTOR 0..4
    else if (a == "--particle") { need(1); particle = argv[++i]; }            // proton | neutron
    else if (a == "--Nstoch") { need(1); Nstoch = atoi(argv[++i]); }           // the loop flags of qkxtm/QKXTM_util.cpp (Calc_Loops.cpp:87-103)
    else if (a == "--NdumpStep") { need(1); NdumpStep = atoi(argv[++i]); }
    else if (a == "--k-probing") { need(1); k_probing = atoi(argv[++i]); }
    else if (a == "--hadamLow") { need(1); hadamLow = atoi(argv[++i]); }
    else if (a == "--hadamHigh") { need(1); hadamHigh = atoi(argv[++i]); }
    else if (a == "--spinColorDil") { need(1); spinColorDil = std::string(argv[++i]) == "yes"; }
    else if (a == "--isFullOp") { need(1); isFullOp = std::string(argv[++i]) == "yes"; }
    else if (a == "--source-type") { need(1); source_type = argv[++i]; }      // random | unity
    else if (a == "--defl-steps") { need(1); n_defl_steps = atoi(argv[++i]); need(n_defl_steps); for (int d = 0; d < n_defl_steps && d < MAX_DEFLSTEPS; d++) defl_step_nEv[d] = atoi(argv[++i]); }
    else if (a == "--nsrc") { need(1); nsrc = atoi(argv[++i]); }               // columns of the pipelined multi-RHS leg of --test e2e
    else if (a == "--e2e-reps") { need(1); e2e_reps = atoi(argv[++i]); }
    else if (a == "--help") { usage(); return 0; }
    else { fprintf(stderr, "unknown flag %s\n", a.c_str()); usage(); return 2; }
  }
  const long long V = (long long)dim[0] * dim[1] * dim[2] * dim[3];
  int coord[4] = {0, 0, 0, 0};

  // setGaugeParam (qkxtm/Calc_Loops.cpp:189-225)
  QudaGaugeParam gauge_param = newQudaGaugeParam();
  for (int d = 0; d < 4; d++) gauge_param.X[d] = dim[d];
  gauge_param.anisotropy = 1.0;
  gauge_param.type = QUDA_WILSON_LINKS;
  gauge_param.gauge_order = QUDA_QDP_GAUGE_ORDER;
  gauge_param.t_boundary = QUDA_ANTI_PERIODIC_T;
  gauge_param.cpu_prec = QUDA_DOUBLE_PRECISION;
  gauge_param.cuda_prec = QUDA_DOUBLE_PRECISION;
  gauge_param.cuda_prec_sloppy = prec_sloppy == "single" ? QUDA_SINGLE_PRECISION : QUDA_DOUBLE_PRECISION;
  gauge_param.reconstruct = gauge_param.reconstruct_sloppy = recon == 8 ? QUDA_RECONSTRUCT_8 : (recon == 12 ? QUDA_RECONSTRUCT_12 : QUDA_RECONSTRUCT_NO);
  gauge_param.gauge_fix = QUDA_GAUGE_FIXED_NO;
  gauge_param.ga_pad = 0;

  // setInvertParam (qkxtm/Calc_Loops.cpp:380-497)
  QudaInvertParam inv_param = newQudaInvertParam();
  inv_param.Ls = 1; inv_param.sp_pad = 0; inv_param.cl_pad = 0;
  inv_param.kappa = kappa < 0 ? 1.0 / (2.0 * (1 + 3 / gauge_param.anisotropy + mass)) : kappa;   // Calc_Loops.cpp:382-388
  inv_param.mass = 0.5 / inv_param.kappa - (1 + 3 / gauge_param.anisotropy);
  inv_param.mu = mu;
  inv_param.dslash_type = dslash_type == "twisted-clover" ? QUDA_TWISTED_CLOVER_DSLASH : QUDA_TWISTED_MASS_DSLASH;
  inv_param.clover_coeff = csw * inv_param.kappa;                              // qkxtm/MG_Bench.cpp:249
  inv_param.twist_flavor = QUDA_TWIST_SINGLET;
  inv_param.cpu_prec = inv_param.cuda_prec = QUDA_DOUBLE_PRECISION;
  inv_param.cuda_prec_sloppy = inv_param.cuda_prec_precondition = gauge_param.cuda_prec_sloppy;
  inv_param.preserve_source = QUDA_PRESERVE_SOURCE_NO;
  inv_param.gamma_basis = QUDA_UKQCD_GAMMA_BASIS;
  inv_param.dirac_order = QUDA_DIRAC_ORDER;
  inv_param.input_location = inv_param.output_location = QUDA_CPU_FIELD_LOCATION;
  inv_param.solution_type = QUDA_MAT_SOLUTION;
  inv_param.solve_type = QUDA_NORMOP_PC_SOLVE;
  inv_param.matpc_type = matpc == "odd-odd" ? QUDA_MATPC_ODD_ODD : (matpc == "even-even-asym" ? QUDA_MATPC_EVEN_EVEN_ASYMMETRIC
                         : (matpc == "odd-odd-asym" ? QUDA_MATPC_ODD_ODD_ASYMMETRIC : QUDA_MATPC_EVEN_EVEN));
  inv_param.inv_type = QUDA_CG_INVERTER;
  inv_param.mass_normalization = massnorm == "mass" ? QUDA_MASS_NORMALIZATION : QUDA_KAPPA_NORMALIZATION;
  inv_param.residual_type = QUDA_L2_RELATIVE_RESIDUAL;
  inv_param.tol = tol; inv_param.maxiter = niter;
  inv_param.reliable_delta = 1e-4;                                             // qkxtm/Calc_Loops.cpp:481
  inv_param.verbosity = verb == "silent" ? QUDA_SILENT : (verb == "verbose" ? QUDA_VERBOSE : QUDA_SUMMARIZE);
  setVerbosityQuda(inv_param.verbosity);

  qudaQKXTMinfo info;
  memset(&info, 0, sizeof(info));
  for (int d = 0; d < 4; d++) info.lL[d] = dim[d];
  info.isEven = ((int)inv_param.matpc_type & 1) == 0;
  info.nsmearGauss = nsmearGauss; info.alphaGauss = alphaGauss;
  info.kappa = inv_param.kappa; info.mu = mu; info.inv_tol = tol; info.Precision = QUDA_DOUBLE_PRECISION;
  info.Nsources = 1; info.Q_sq = q_sq; info.CorrSpace = MOMENTUM_SPACE; info.CorrFileFormat = ASCII_FORM;
  for (int d = 0; d < 4; d++) info.sourcePosition[0][d] = src_pos[d];
  if (tsink > 0) { info.run3pt_src[0] = 1; info.Ntsink = 1; info.tsinkSource[0] = tsink; info.Nproj[0] = 1; info.proj_list[0][0] = proj; }

  // one process per rank (torchrun --no-python / mpirun / srun export the rank): qkxtm/QKXTM_util.cpp:48-68
  initCommsGridQuda(4, grid, nullptr, nullptr);
  for (int d = 0; d < 4; d++) coord[d] = comm_coord(d);
  const bool root = comm_rank() == 0;
  if (comm_size() > 1 && !out.empty()) out += ".rank" + std::to_string(comm_rank());

  // synthetic configuration: random SU(3), QDP even-odd order, anti-periodic T folded in (this rank's block of the global field)
  std::vector<double> gbuf((size_t)4 * V * 18);
  double *gauge[4] = {gbuf.data(), gbuf.data() + (size_t)V * 18, gbuf.data() + (size_t)2 * V * 18, gbuf.data() + (size_t)3 * V * 18};
  tmq_fieldgen_gauge_qdp(gauge, dim, grid, coord, 137, -1);

  initQuda(comm_size() > 1 ? -1 : 0);
  init_qudaQKXTM(&info);
  printf_qudaQKXTM();
  loadGaugeQuda((void *)gauge, &gauge_param);
  if (inv_param.dslash_type == QUDA_TWISTED_CLOVER_DSLASH) {                    // qkxtm/MG_Bench.cpp:605-608
    printf("Constructing clover field\n");
    loadCloverQuda(NULL, NULL, &inv_param);
    printf("Clover field done\n");
  }

  std::vector<double> result;
  if (test == "invert" || test == "mat") {
    std::vector<double> b((size_t)V * 24), x((size_t)V * 24);
    if (source == "z4") tmq_fieldgen_spinor_z4(b.data(), dim, grid, coord, seed, 1);
    else tmq_fieldgen_spinor_gaussian(b.data(), dim, grid, coord, seed, 1);
    if (test == "invert") invertQuda(x.data(), b.data(), &inv_param);
    else MatQuda(x.data(), b.data(), &inv_param);
    result = x;
  } else if (test == "loops" || test == "mdagm") {
    std::vector<double> b((size_t)V * 24), x((size_t)V * 24);
    if (source == "z4") tmq_fieldgen_spinor_z4(b.data(), dim, grid, coord, seed, 0);     // plug-in host order [x_lex][s][c]
    else tmq_fieldgen_spinor_gaussian(b.data(), dim, grid, coord, seed, 0);
    if (test == "loops") calc_loops_solve(x.data(), b.data(), &inv_param, info);
    else ApplyMdagM(x.data(), b.data(), &inv_param, info.isEven);
    result = x;
  } else if (test == "eig") {
    // calcEigenVectors-style flow (lib/qudaQKXTM_interface.cpp:1378-1395): eigenSolver, then deflate a source with it
    qudaQKXTM_arpackInfo ai;
    memset(&ai, 0, sizeof(ai));
    ai.PolyDeg = polydeg; ai.nEv = nev; ai.nKv = nkv; ai.spectrumPart = SR; ai.isACC = polydeg > 0;
    ai.tolArpack = eig_tol; ai.maxIterArpack = 1000; ai.amin = amin; ai.amax = amax; ai.isEven = info.isEven; ai.isFullOp = false;
    QKXTM_Deflation<double> *deflation = new QKXTM_Deflation<double>(&inv_param, ai);
    deflation->printInfo();
    deflation->eigenSolver();
    std::vector<double> b((size_t)V * 24), v0((size_t)V * 24);
    if (source == "z4") tmq_fieldgen_spinor_z4(b.data(), dim, grid, coord, seed, 0);
    else tmq_fieldgen_spinor_gaussian(b.data(), dim, grid, coord, seed, 0);
    QKXTM_Vector<double> K_in(BOTH, VECTOR), K_defl(BOTH, VECTOR);
    memcpy(K_in.H_elem(), b.data(), b.size() * sizeof(double));          // deflateVector reads the host AoS vector
    deflation->deflateVector(K_defl, K_in);
    K_defl.download();
    deflation->copyEigenVectorToQKXTM_Vector(0, v0.data());
    for (int i = 0; i < nev; i++) result.push_back(deflation->EigenValues()[2 * i]);
    for (int i = 0; i < nev; i++) result.push_back(deflation->Residuals()[i]);
    result.insert(result.end(), v0.begin(), v0.end());
    result.insert(result.end(), K_defl.H_elem(), K_defl.H_elem() + (size_t)V * 24);
    inv_param.iter = deflation->MatVecs();
    delete deflation;
  } else if (test == "e2e") {
    // END-TO-END timing through the reference-facing calls with HOST buffers (bench.py's e2e leg): page-locked host sources and
    // solutions, every copy inside the timed region.  (1) one invertQuda, fp64; (2) one invertQuda with the drivers' usual fp32 sloppy
    // precision; (3) nsrc columns through invertMultiSrcQuda, fp64, uploads / downloads of neighbouring columns behind each solve.
    const size_t nbytes = (size_t)V * 24 * sizeof(double);
    double *hb[2], *hx[2];
    for (int k = 0; k < 2; k++) {
      void *p = nullptr;
      if (tmq_host_alloc_pinned(qkxtm_context(), &p, nbytes)) { fprintf(stderr, "%s\n", tmq_last_error()); return 1; }
      hb[k] = (double *)p;
      if (tmq_host_alloc_pinned(qkxtm_context(), &p, nbytes)) { fprintf(stderr, "%s\n", tmq_last_error()); return 1; }
      hx[k] = (double *)p;
      tmq_fieldgen_spinor_z4(hb[k], dim, grid, coord, seed + k, 1);
    }
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const QudaPrecision sloppy_saved = inv_param.cuda_prec_sloppy;
    inv_param.cuda_prec_sloppy = QUDA_DOUBLE_PRECISION;
    invertQuda(hx[0], hb[0], &inv_param);                                       // warm-up (module load, scratch allocation)
    comm_barrier();
    auto t0 = now();
    for (int r = 0; r < e2e_reps; r++) invertQuda(hx[r & 1], hb[r & 1], &inv_param);
    comm_barrier();
    const double t_single = secs(t0, now()) / e2e_reps;
    const int it_single = inv_param.iter;
    const double res_single = inv_param.true_res;
    inv_param.cuda_prec_sloppy = QUDA_SINGLE_PRECISION;
    invertQuda(hx[0], hb[0], &inv_param);
    comm_barrier();
    t0 = now();
    for (int r = 0; r < e2e_reps; r++) invertQuda(hx[r & 1], hb[r & 1], &inv_param);
    comm_barrier();
    const double t_mixed = secs(t0, now()) / e2e_reps;
    const int it_mixed = inv_param.iter;
    const double res_mixed = inv_param.true_res;
    inv_param.cuda_prec_sloppy = QUDA_DOUBLE_PRECISION;
    std::vector<void *> pb(nsrc), px(nsrc);
    for (int k = 0; k < nsrc; k++) { pb[k] = hb[k & 1]; px[k] = hx[k & 1]; }
    inv_param.num_src = nsrc;
    comm_barrier();
    t0 = now();
    invertMultiSrcQuda(px.data(), pb.data(), &inv_param);
    comm_barrier();
    const double t_multi = secs(t0, now());
    const int it_multi = inv_param.iter;
    const double res_multi = inv_param.true_res, solver_secs_multi = inv_param.secs;
    double last_loop_secs = 0;                                                  // iteration loop of the last column's CG, and the ghost exchange in use
    int last_reliable = 0;
    tmq_cg_stats(qkxtm_context(), &last_loop_secs, &last_reliable);
    const int halo_mode = tmq_halo_mode(qkxtm_context());
    inv_param.cuda_prec_sloppy = sloppy_saved;
    // the last solution is returned (even-odd host order) so that the caller can check it
    result.assign(hx[(nsrc - 1) & 1], hx[(nsrc - 1) & 1] + (size_t)V * 24);
    // the host link these numbers are bounded by, with every rank of the job copying at once: uploads alone, downloads alone, both
    double link[3] = {0, 0, 0};
    {
      const int reps = 3;
      double s3[3];
      comm_barrier();
      if (tmq_host_link_probe(qkxtm_context(), hb[0], hx[nsrc & 1], reps, s3)) { fprintf(stderr, "%s\n", tmq_last_error()); return 1; }
      for (int i = 0; i < 3; i++) link[i] = (double)nbytes * reps * (i == 2 ? 2 : 1) / s3[i] * 1e-9;      // this rank's GB/s
      if (tmq_allreduce_host(qkxtm_context(), link, 3)) { fprintf(stderr, "%s\n", tmq_last_error()); return 1; }   // summed over the ranks
    }
    if (root)
      printf("RESULT_E2E {\"single_secs\": %.6f, \"single_iter\": %d, \"single_true_res\": %.6e, \"mixed_secs\": %.6f, \"mixed_iter\": %d, "
             "\"mixed_true_res\": %.6e, \"nsrc\": %d, \"multi_secs\": %.6f, \"multi_iter\": %d, \"multi_true_res\": %.6e, \"multi_solver_secs\": %.6f, "
             "\"bytes_per_field\": %zu, \"link_h2d_gbs\": %.2f, \"link_d2h_gbs\": %.2f, \"link_duplex_gbs\": %.2f, \"halo_mode\": %d, \"last_column_loop_secs\": %.6f}\n", t_single, it_single, res_single, t_mixed, it_mixed,
             res_mixed, nsrc, t_multi, it_multi, res_multi, solver_secs_multi, nbytes, link[0], link[1], link[2], halo_mode, last_loop_secs);
    for (int k = 0; k < 2; k++) { tmq_host_free_pinned(qkxtm_context(), hb[k]); tmq_host_free_pinned(qkxtm_context(), hx[k]); }
  } else if (test == "calcloops") {
    // qkxtm/Calc_Loops.cpp main() (:585-791): arpackInfo, loopInfo, the operator of the eigensolver, then calc_loops.  The hook stands where
    // the reference contracts: it records every eigenvalue of the exact part and, for every solve and deflation step, the host source and
    // the projected solution (both in the plug-in's AoS order), so that the test can check the whole chain against the CPU oracle.
    qudaQKXTM_arpackInfo arpackInfo;
    memset(&arpackInfo, 0, sizeof(arpackInfo));
    arpackInfo.PolyDeg = polydeg; arpackInfo.nEv = nev; arpackInfo.nKv = nkv; arpackInfo.isACC = polydeg > 0; arpackInfo.tolArpack = eig_tol;
    arpackInfo.maxIterArpack = 1000; arpackInfo.amin = amin; arpackInfo.amax = amax; arpackInfo.isEven = info.isEven; arpackInfo.isFullOp = isFullOp;
    arpackInfo.spectrumPart = SR;
    info.source_type = source_type == "unity" ? UNITY : RANDOM;
    qudaQKXTM_loopInfo loopInfo;
    memset((void *)&loopInfo, 0, sizeof(loopInfo));
    loopInfo.Nstoch = Nstoch; loopInfo.seed = (unsigned long int)seed; loopInfo.Ndump = NdumpStep; loopInfo.traj = 0; loopInfo.Qsq = q_sq;
    loopInfo.k_probing = k_probing; loopInfo.spinColorDil = spinColorDil; loopInfo.hadamLow = hadamLow; loopInfo.hadamHigh = hadamHigh;
    snprintf(loopInfo.loop_fname, sizeof(loopInfo.loop_fname), "%s", out.empty() ? "loop" : out.c_str());
    loopInfo.kappa = inv_param.kappa; loopInfo.csw = csw; loopInfo.mu = mu; loopInfo.inv_tol = tol; loopInfo.FileFormat = ASCII_FORM;
    loopInfo.Nprint = loopInfo.Nstoch / loopInfo.Ndump;
    if (n_defl_steps == 0) { loopInfo.nSteps_defl = 1; loopInfo.deflStep[0] = nev; }       // Calc_Loops.cpp:656-659
    else { loopInfo.nSteps_defl = n_defl_steps; for (int a_ = 0; a_ < n_defl_steps; a_++) loopInfo.deflStep[a_] = defl_step_nEv[a_]; }
    if (inv_param.mu > 0) inv_param.mu = -inv_param.mu;                        // "For the loops we invert the negative mu" (Calc_Loops.cpp:424)
    QudaInvertParam EVinv_param = inv_param;                                    // Calc_Loops.cpp:709-715
    EVinv_param.matpc_type = info.isEven ? QUDA_MATPC_EVEN_EVEN_ASYMMETRIC : QUDA_MATPC_ODD_ODD_ASYMMETRIC;
    EVinv_param.mass_normalization = QUDA_MASS_NORMALIZATION;
    std::vector<double> lex((size_t)4 * V * 18);
    double *glex[4];
    for (int mu_ = 0; mu_ < 4; mu_++) {
      glex[mu_] = lex.data() + (size_t)mu_ * V * 18;
      for (long long i = 0; i < V; i++) {
        const int x0 = i % dim[0], y = (i / dim[0]) % dim[1], z = (i / ((long long)dim[0] * dim[1])) % dim[2], t = i / ((long long)dim[0] * dim[1] * dim[2]);
        const long long eo = (long long)((x0 + y + z + t) & 1) * (V / 2) + i / 2;
        memcpy(glex[mu_] + i * 18, gauge[mu_] + eo * 18, 18 * sizeof(double));
      }
    }
    struct Rec { std::vector<double> *out; long long V; bool isEven; } rec = {&result, V, info.isEven};
    qkxtm_set_loop_hook([](const qkxtm_loop_event *ev, void *user) {
      Rec *r = (Rec *)user;
      QKXTM_Vector<double> K(BOTH, VECTOR);
      K.downloadFromCuda(ev->x, r->isEven);
      K.download();
      const double head[8] = {(double)ev->kind, (double)ev->is, (double)ev->ih, (double)ev->sc, (double)ev->dstep, (double)ev->NeV_defl,
                              ev->kind == 0 ? ev->eigenvalue : (double)ev->iter, ev->true_res};
      r->out->insert(r->out->end(), head, head + 8);
      if (ev->kind == 1) r->out->insert(r->out->end(), ev->h_source, ev->h_source + (size_t)r->V * 24);
      r->out->insert(r->out->end(), K.H_elem(), K.H_elem() + (size_t)r->V * 24);
    }, &rec);
    calc_loops((void **)glex, &EVinv_param, &inv_param, &gauge_param, arpackInfo, loopInfo, info);
    qkxtm_set_loop_hook(nullptr, nullptr);
  } else if (test == "lowmodes") {
    // qkxtm/CalcLowModeProjection.cpp main(): the low modes of the asymmetric even-odd M^dag M
    qudaQKXTM_arpackInfo ai;
    memset(&ai, 0, sizeof(ai));
    ai.PolyDeg = polydeg; ai.nEv = nev; ai.nKv = nkv; ai.spectrumPart = SR; ai.isACC = polydeg > 0;
    ai.tolArpack = eig_tol; ai.maxIterArpack = 1000; ai.amin = amin; ai.amax = amax; ai.isEven = info.isEven; ai.isFullOp = false;
    int nconv = 0;
    result.resize(nev);
    calcLowModeProjection(&inv_param, ai, &nconv, result.data());
    inv_param.iter = nconv;
  } else if (test == "mgbench" || test == "twop") {
    // lexicographic copy of the links for the plaquette print (gauge_Plaq in the drivers): unit test uses the
    // same synthetic field reordered even-odd -> lexicographic
    std::vector<double> lex((size_t)4 * V * 18);
    double *glex[4];
    for (int mu_ = 0; mu_ < 4; mu_++) {
      glex[mu_] = lex.data() + (size_t)mu_ * V * 18;
      for (long long i = 0; i < V; i++) {
        const int x0 = i % dim[0], y = (i / dim[0]) % dim[1], z = (i / ((long long)dim[0] * dim[1])) % dim[2], t = i / ((long long)dim[0] * dim[1] * dim[2]);
        const long long eo = (long long)((x0 + y + z + t) & 1) * (V / 2) + i / 2;
        memcpy(glex[mu_] + i * 18, gauge[mu_] + eo * 18, 18 * sizeof(double));
      }
    }
    if (test == "mgbench") {
      result.resize((size_t)12 * V * 24);
      if (nsmearGauss > 0) testGaussSmearing((void **)glex);                       // lib/qudaQKXTM_utils.cpp:116-141
      MG_bench((void **)glex, (void **)gauge, &gauge_param, &inv_param, info, result.data());
    } else {
      // qkxtm/CalcMG_2pt3pt_EvenOdd.cpp main(): the two-point function of one source position
      std::string base = out.empty() ? std::string("twop") : out;
      if (comm_size() > 1 && !out.empty()) base = out.substr(0, out.rfind(".rank"));      // correlator files are written once, by rank 0
      std::vector<char> f2(base.begin(), base.end()); f2.push_back(0);
      std::string base3 = base + ".threep";
      std::vector<char> f3(base3.begin(), base3.end()); f3.push_back(0);
      // both link arrays in the lexicographic layout of packGauge; on a process grid the links of the derivative insertions are withheld
      // (gauge = NULL: ultra-local insertion only -- the library refuses the derivative insertions on a split lattice)
      calcMG_threepTwop_EvenOdd((void **)glex, comm_size() > 1 ? (void **)NULL : (void **)glex, &gauge_param, &inv_param, info, f2.data(), f3.data(),
                                particle == "neutron" ? NEUTRON : PROTON);
    }
  } else { usage(); return 2; }

  if (root)
    printf("RESULT test=%s iter=%d true_res=%.6e secs=%.6f gflops=%.3f kappa=%.17g mu=%.17g\n", test.c_str(), inv_param.iter,
           inv_param.true_res, inv_param.secs, inv_param.gflops, inv_param.kappa, inv_param.mu);
  if (!out.empty()) {
    FILE *f = fopen(out.c_str(), "wb");
    if (!f) { perror("fopen"); return 1; }
    fwrite(result.data(), sizeof(double), result.size(), f);
    fclose(f);
  }
  freeGaugeQuda();
  endQuda();
  return 0;
}
