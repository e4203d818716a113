// sqc_mmd_host.inl — C-ABI entry points of the MMD path (included by sqc_lib.cu inside extern "C").  Kernels: sqc_mmd.cuh.
namespace {
void mmd_run(int device, std::vector<MmdJob>& jobs, int D, double sigma, int absolute, int n_times, double* per_job_host, double* mean_host) {
  if (!(sigma > 0)) fail(SQC_ERR_BAD_ARG, "sigma must be > 0");                      // mmd_fn.cpp:28-30
  if (D < 1 || D > SQC_MAXD) fail(SQC_ERR_BAD_ARG, "D must be in 1..%d", SQC_MAXD);
  const int nj = (int)jobs.size();
  std::vector<int> slice_job;
  for (int j = 0; j < nj; ++j) {
    MmdJob& b = jobs[j];
    if (b.nx < 1 || b.ny < 1) fail(SQC_ERR_BAD_ARG, "empty matrix in an MMD job");
    b.slice0 = (int)slice_job.size();
    b.nslices = (std::max(b.nx, b.ny) + MMD_THREADS - 1) / MMD_THREADS;
    for (int s = 0; s < b.nslices; ++s) slice_job.push_back(j);
  }
  const size_t ns = slice_job.size();
  MmdJob* dj = nullptr; int* dsj = nullptr; double* dpart = nullptr; double* dper = nullptr; double* dmean = nullptr;
  auto cleanup = [&] { cudaFree(dj); cudaFree(dsj); cudaFree(dpart); cudaFree(dper); cudaFree(dmean); };
  const int ng = n_times > 0 ? nj / n_times : 0;
  try {
    CK(cudaMalloc((void**)&dj, (size_t)nj * sizeof(MmdJob))); CK(cudaMalloc((void**)&dsj, ns * sizeof(int))); CK(cudaMalloc((void**)&dpart, 3 * ns * sizeof(double)));
    CK(cudaMalloc((void**)&dper, (size_t)nj * sizeof(double))); if (ng) CK(cudaMalloc((void**)&dmean, (size_t)ng * sizeof(double)));
    CK(cudaMemcpy(dj, jobs.data(), (size_t)nj * sizeof(MmdJob), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dsj, slice_job.data(), ns * sizeof(int), cudaMemcpyHostToDevice));
#define SQC_CASE_MMD(DD) case DD: k_mmd_slices<DD><<<(unsigned)ns, MMD_THREADS>>>(dj, dsj, sigma, dpart); break;
    switch (D) { SQC_FOR_D(SQC_CASE_MMD) default: fail(SQC_ERR_BAD_ARG, "D = %d not compiled in", D); }
    const int nthreads = n_times > 0 ? ng : nj;
    k_mmd_finish<<<(nthreads + 127) / 128, 128>>>(dj, nj, dpart, absolute, n_times, dper, dmean);
    CK(cudaGetLastError());
    if (per_job_host) CK(cudaMemcpy(per_job_host, dper, (size_t)nj * sizeof(double), cudaMemcpyDeviceToHost));
    if (mean_host && ng) CK(cudaMemcpy(mean_host, dmean, (size_t)ng * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaDeviceSynchronize());
  } catch (...) { cleanup(); throw; }
  cleanup();
}
}  // namespace

// mmd_fn(X, Y, sigma), reference src/mmd_fn.cpp:20-91: X n_x x D, Y n_y x D column-major (R matrices), *mmd the signed value
int sqc_mmd_fn(const double* X, int64_t n_x, const double* Y, int64_t n_y, int32_t D, double sigma, double* mmd, int32_t device, char* err, size_t errlen) {
  return guarded([&] {
    if (!X || !Y || !mmd) fail(SQC_ERR_BAD_ARG, "null input");
    if (n_x < 1 || n_y < 1 || n_x > 2000000000ll || n_y > 2000000000ll) fail(SQC_ERR_BAD_ARG, "bad dimensions");
    int dev; select_device(device, &dev);
    double *dx = nullptr, *dy = nullptr;
    try {
      CK(cudaMalloc((void**)&dx, (size_t)n_x * D * sizeof(double))); CK(cudaMalloc((void**)&dy, (size_t)n_y * D * sizeof(double)));
      CK(cudaMemcpy(dx, X, (size_t)n_x * D * sizeof(double), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dy, Y, (size_t)n_y * D * sizeof(double), cudaMemcpyHostToDevice));
      std::vector<MmdJob> jobs(1);
      jobs[0] = MmdJob{dx, dy, nullptr, nullptr, (long long)n_x, (long long)n_y, (int)n_x, (int)n_y, 0, 0};
      mmd_run(dev, jobs, D, sigma, 0, 0, mmd, nullptr);
    } catch (...) { cudaFree(dx); cudaFree(dy); throw; }
    cudaFree(dx); cudaFree(dy);
  }, err, errlen);
}

// .calc_mmd_mat (reference R/calc_pairwise_mmds.R:198-247) in one call: for every pair (i, j) and every one of n_times
// repetitions, |mmd_fn| of the rows .dist_fn's sample() drew (:265-280), then the mean over the repetitions (:226-231)
int sqc_mmd_pairs(const sqc_mmd_args* a, double* mmd_mean, double* mmd_all, char* err, size_t errlen) {
  return guarded([&] {
    if (!a || !a->mats || !a->mat_rows || !a->pair_i || !a->pair_j || !mmd_mean) fail(SQC_ERR_BAD_ARG, "null input");
    if (a->abi_version != SQC_ABI_VERSION) fail(SQC_ERR_BAD_ARG, "ABI version mismatch");
    if (a->n_samples < 1 || a->n_pairs < 1 || a->n_times < 1 || a->subsample < 1) fail(SQC_ERR_BAD_ARG, "bad dimensions");
    if ((double)a->n_pairs * a->n_times > 2.0e8) fail(SQC_ERR_BAD_ARG, "too many (pair, repetition) jobs for one call: split the pair list");
    int dev; select_device(a->device, &dev);
    const int D = a->D, S = a->n_samples, T = a->n_times, sub = a->subsample;
    std::vector<size_t> off((size_t)S + 1, 0);
    for (int s = 0; s < S; ++s) { if (a->mat_rows[s] < 1) fail(SQC_ERR_BAD_ARG, "sample %d has no rows", s); off[(size_t)s + 1] = off[s] + (size_t)a->mat_rows[s] * D; }
    bool need_idx = false;
    for (long long p = 0; p < a->n_pairs; ++p) {
      const int i = a->pair_i[p], j = a->pair_j[p];
      if (i < 0 || i >= S || j < 0 || j >= S) fail(SQC_ERR_BAD_ARG, "pair %lld out of range", p);
      if (a->mat_rows[i] > sub || a->mat_rows[j] > sub) need_idx = true;
    }
    if (need_idx && (!a->idx_i || !a->idx_j)) fail(SQC_ERR_BAD_ARG, "subsample indices missing");
    double* dm = nullptr; int *di = nullptr, *dj2 = nullptr;
    const size_t nidx = (size_t)a->n_pairs * T * sub;
    try {
      CK(cudaMalloc((void**)&dm, off[S] * sizeof(double)));
      CK(cudaMemcpy(dm, a->mats, off[S] * sizeof(double), cudaMemcpyHostToDevice));
      if (need_idx) {
        CK(cudaMalloc((void**)&di, nidx * sizeof(int))); CK(cudaMalloc((void**)&dj2, nidx * sizeof(int)));
        CK(cudaMemcpy(di, a->idx_i, nidx * sizeof(int), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dj2, a->idx_j, nidx * sizeof(int), cudaMemcpyHostToDevice));
        for (long long p = 0; p < a->n_pairs; ++p) for (int t = 0; t < T; ++t) for (int q = 0; q < sub; ++q) {
          const size_t e = ((size_t)p * T + t) * sub + q;
          if (a->mat_rows[a->pair_i[p]] > sub && (a->idx_i[e] < 0 || a->idx_i[e] >= a->mat_rows[a->pair_i[p]])) fail(SQC_ERR_BAD_ARG, "subsample index out of range");
          if (a->mat_rows[a->pair_j[p]] > sub && (a->idx_j[e] < 0 || a->idx_j[e] >= a->mat_rows[a->pair_j[p]])) fail(SQC_ERR_BAD_ARG, "subsample index out of range");
        }
      }
      std::vector<MmdJob> jobs((size_t)a->n_pairs * T);
      for (long long p = 0; p < a->n_pairs; ++p) for (int t = 0; t < T; ++t) {
        const int i = a->pair_i[p], j = a->pair_j[p];
        const bool si = a->mat_rows[i] > sub, sj = a->mat_rows[j] > sub;                 // if (n_i > subsample), :271, :276
        const size_t e = ((size_t)p * T + t) * sub;
        jobs[(size_t)p * T + t] = MmdJob{dm + off[i], dm + off[j], si ? di + e : nullptr, sj ? dj2 + e : nullptr, (long long)a->mat_rows[i], (long long)a->mat_rows[j],
                                         si ? sub : (int)a->mat_rows[i], sj ? sub : (int)a->mat_rows[j], 0, 0};
      }
      mmd_run(dev, jobs, D, a->sigma, 1, T, mmd_all, mmd_mean);
    } catch (...) { cudaFree(dm); cudaFree(di); cudaFree(dj2); throw; }
    cudaFree(dm); cudaFree(di); cudaFree(dj2);
  }, err, errlen);
}
