// sqc_init_host.inl — C-ABI entry point of the device .init_clusts pieces (included by sqc_lib.cu inside extern "C").
// Kernels: sqc_init.cuh.
// .init_clusts on the device (R/fit_sampleqc.R:277-331): per-sample medians, then (optionally) the classification of
// all cells under the mixture the caller fitted on its subsample
int sqc_init_clusts(const sqc_init_args* a, double* medians, double* gamma, int32_t* z, char* err, size_t errlen) {
  return guarded([&] {
    if (!a || !a->x || !a->groups || !medians) fail(SQC_ERR_BAD_ARG, "null input");
    if (a->abi_version != SQC_ABI_VERSION) fail(SQC_ERR_BAD_ARG, "ABI version mismatch");
    const int D = a->D, J = a->J; const long long N = a->N;
    const bool classify = a->pro != nullptr;
    const int K = classify ? a->K : 0;
    if (D < 1 || D > SQC_MAXD || J < 1 || N < 1 || (classify && (K < 1 || K > SQC_MAXK || !a->mean || !a->sigma))) fail(SQC_ERR_BAD_ARG, "bad dimensions");
    if ((gamma || z) && !classify) fail(SQC_ERR_BAD_ARG, "gamma / z requested without a mixture");
    int dev; select_device(a->device, &dev);
    // segment bounds; cells of a sample need not be contiguous (stable counting sort by sample if they are not)
    std::vector<long long> bounds((size_t)J + 1, 0);
    bool sorted = true;
    for (long long i = 0; i < N; ++i) {
      const int g = a->groups[i];
      if (g < 0 || g >= J) fail(SQC_ERR_BAD_ARG, "groups out of range");
      bounds[(size_t)g + 1]++;
      if (i && g < a->groups[i - 1]) sorted = false;
    }
    for (int j = 0; j < J; ++j) bounds[(size_t)j + 1] += bounds[j];
    // mixture parameters: mean, inverse Cholesky factor, log-density constant
    const int P = D + D * D + 1;
    std::vector<double> par((size_t)std::max(K, 1) * P, 0.0);
    for (int k = 0; k < K; ++k) {
      double L[SQC_MAXD * SQC_MAXD], Li[SQC_MAXD * SQC_MAXD];
      if (!small_chol_inv(a->sigma + (size_t)k * D * D, D, L, Li)) fail(SQC_ERR_NOT_SPD, "mixture covariance %d is not positive definite", k);
      double c = log(a->pro[k]) - 0.5 * D * 1.8378770664093454835606594728112;
      for (int d = 0; d < D; ++d) { par[(size_t)k * P + d] = a->mean[k + (size_t)K * d]; c -= log(L[d + D * d]); }
      for (int e = 0; e < D * D; ++e) par[(size_t)k * P + D + e] = Li[e];
      par[(size_t)k * P + D + D * D] = c;
    }
    double *dx = nullptr, *dmed = nullptr, *dpar = nullptr, *dgam = nullptr; int *dg = nullptr, *dz = nullptr, *dflag = nullptr; long long* db = nullptr;
    auto cleanup = [&] { cudaFree(dx); cudaFree(dmed); cudaFree(dpar); cudaFree(dgam); cudaFree(dg); cudaFree(dz); cudaFree(dflag); cudaFree(db); };
    try {
      CK(cudaMalloc((void**)&dx, (size_t)N * D * sizeof(double))); CK(cudaMalloc((void**)&dmed, (size_t)J * D * sizeof(double)));
      CK(cudaMalloc((void**)&db, ((size_t)J + 1) * sizeof(long long))); CK(cudaMalloc((void**)&dflag, sizeof(int)));
      CK(cudaMemset(dflag, 0, sizeof(int)));
      CK(cudaMemcpy(db, bounds.data(), ((size_t)J + 1) * sizeof(long long), cudaMemcpyHostToDevice));
      std::vector<long long> pos;                               // pos[i]: where cell i sits in sample order
      if (sorted) CK(cudaMemcpy(dx, a->x, (size_t)N * D * sizeof(double), cudaMemcpyHostToDevice));
      else {
        std::vector<long long> next(bounds.begin(), bounds.end() - 1);
        pos.resize((size_t)N);
        for (long long i = 0; i < N; ++i) pos[(size_t)i] = next[(size_t)a->groups[i]]++;
        std::vector<double> col((size_t)N);
        for (int d = 0; d < D; ++d) {
          for (long long i = 0; i < N; ++i) col[(size_t)pos[(size_t)i]] = a->x[(size_t)d * N + i];
          CK(cudaMemcpy(dx + (size_t)d * N, col.data(), (size_t)N * sizeof(double), cudaMemcpyHostToDevice));
        }
      }
      k_sample_median<<<J * D, MED_THREADS>>>(dx, db, N, J, dmed, dflag);
      CK(cudaGetLastError());
      int flag = 0;
      CK(cudaMemcpy(&flag, dflag, sizeof(int), cudaMemcpyDeviceToHost));
      if (flag) fail(SQC_ERR_NAN, "NaN in x");
      CK(cudaMemcpy(medians, dmed, (size_t)J * D * sizeof(double), cudaMemcpyDeviceToHost));
      if (classify && (gamma || z)) {
        CK(cudaMalloc((void**)&dg, (size_t)N * sizeof(int))); CK(cudaMalloc((void**)&dpar, par.size() * sizeof(double)));
        if (gamma) CK(cudaMalloc((void**)&dgam, (size_t)N * K * sizeof(double)));
        if (z) CK(cudaMalloc((void**)&dz, (size_t)N * sizeof(int)));
        if (!sorted) CK(cudaMemcpy(dx, a->x, (size_t)N * D * sizeof(double), cudaMemcpyHostToDevice));   // classify in the caller's order
        CK(cudaMemcpy(dg, a->groups, (size_t)N * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dpar, par.data(), par.size() * sizeof(double), cudaMemcpyHostToDevice));
        const int grid = (int)std::min<long long>((N + 255) / 256, 148 * 16);
        const size_t sm = (size_t)K * P * sizeof(double);
#define SQC_CASE_CLASSIFY(DD) case DD: k_gmm_classify<DD><<<grid, 256, sm>>>(dx, dg, N, J, K, dmed, dpar, dgam, dz); break;
        switch (D) { SQC_FOR_D(SQC_CASE_CLASSIFY) default: fail(SQC_ERR_BAD_ARG, "D = %d not compiled in", D); }
        CK(cudaGetLastError());
        if (gamma) CK(cudaMemcpy(gamma, dgam, (size_t)N * K * sizeof(double), cudaMemcpyDeviceToHost));
        if (z) CK(cudaMemcpy(z, dz, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost));
      }
      CK(cudaDeviceSynchronize());
    } catch (...) { cleanup(); throw; }
    cleanup();
  }, err, errlen);
}
