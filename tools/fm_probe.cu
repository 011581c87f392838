// tools/fm_probe.cu -- fastmatch_kernel launched on its own (no library context, no replay loop around it), phase
// trace printed: is the kernel slow by itself or in its environment?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -lineinfo -o tools/build/fm_probe tools/fm_probe.cu -ldl
#include "../hardware-acceleration-of-lidar-slam_b200/csrc/score.cu"
int b200slam_set_error(b200slam_ctx *, int code, const char *, ...) { return code; }
extern "C" float b200slam_lattice_value(float p, float s, int k, int n) { return p + (float)(k - n / 2) * s; }
int main()
{
    const int rows = 400, cols = 400, pitch = 416, nb = 1079;
    float *field, *sx, *sy, *hits;
    MatchDev *md;
    MatchHost *mh;
    cudaMalloc(&field, sizeof(float) * (pitch * (rows + 9)));
    cudaMemset(field, 0, sizeof(float) * (pitch * (rows + 9)));
    cudaMalloc(&sx, 4 * 2048); cudaMalloc(&sy, 4 * 2048); cudaMalloc(&hits, 4 * 8192);
    static float hx[2048], hy[2048];
    for (int i = 0; i < nb; ++i) { hx[i] = 5.0f * cosf(0.005f * i); hy[i] = 5.0f * sinf(0.005f * i); }
    cudaMemcpy(sx, hx, 4 * nb, cudaMemcpyHostToDevice); cudaMemcpy(sy, hy, 4 * nb, cudaMemcpyHostToDevice);
    cudaMalloc(&md, sizeof(MatchDev)); cudaMemset(md, 0, sizeof(MatchDev));
    cudaHostAlloc(&mh, sizeof(MatchHost), cudaHostAllocMapped); memset(mh, 0, sizeof(MatchHost));
    FmArgs A = {};
    for (int p = 0; p < 2; ++p) { A.map[p].field = field + 9 * pitch; A.map[p].pitch = pitch; A.map[p].rows = rows; A.map[p].cols = cols; A.map[p].ipixel = 10.0f; }
    A.npass = 1; A.seeded0 = 0; A.scan_x = sx; A.scan_y = sy; A.nbeams = nb; A.nbeams_dev = nullptr;
    A.nth = A.ntx = A.nty = 3; A.match = md; A.hit_values = hits; A.host_result = mh; A.ranges = nullptr; A.trace = 1;
    A.nbp = fastmatch_row_pitch(nb);
    LatticeTables T = {};
    for (int k = 0; k < 3; ++k) { T.v[k] = cosf(0.01f * (k - 1)); T.v[3 + k] = sinf(0.01f * (k - 1)); T.v[6 + k] = 200.0f + 0.5f * k; T.v[9 + k] = 200.0f + 0.5f * k; }
    for (int s1 = 0; s1 < 3; ++s1) for (int k = 0; k < 3; ++k) {
        T.v[FM_TAB_B + 3 * s1 + k] = cosf(0.005f * (k - 1)); T.v[FM_TAB_B + 9 + 3 * s1 + k] = sinf(0.005f * (k - 1));
        T.v[FM_TAB_B + 18 + 3 * s1 + k] = 200.0f + 0.25f * k; T.v[FM_TAB_B + 27 + 3 * s1 + k] = 200.0f + 0.25f * k; }
    const size_t smem = fastmatch_smem_bytes(27, 3, nb);
    cudaFuncSetAttribute(fastmatch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(3); cfg.blockDim = dim3(FM_THREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 3; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    for (int rep = 0; rep < 4; ++rep) {
        A.host_seq = rep + 1;
        cudaLaunchKernelEx(&cfg, fastmatch_kernel, A, T);
        cudaError_t e = cudaDeviceSynchronize();
        printf("rep %d (%s): phases", rep, cudaGetErrorString(e));
        for (int i = 1; i < 9 && mh->trace[i]; ++i) printf(" %lld", mh->trace[i] - mh->trace[i - 1]);
        printf("  | gather end %lld after trace[2]\n", mh->trace[12] - mh->trace[2]);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int np = 1; np <= 2; ++np) for (int tr = 0; tr < 2; ++tr) {
        A.npass = np; A.trace = tr;
        for (int i = 0; i < 20; ++i) cudaLaunchKernelEx(&cfg, fastmatch_kernel, A, T);
        cudaEventRecord(e0);
        for (int i = 0; i < 200; ++i) cudaLaunchKernelEx(&cfg, fastmatch_kernel, A, T);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("npass %d trace %d: %.2f us per launch (%s)\n", np, tr, ms * 5.0f, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
