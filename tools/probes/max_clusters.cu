// How many thread-block clusters of size 1/2/4/8 (one 192-thread CTA with ~225 KB of shared memory per SM, the
// scoring kernel's footprint) can be co-resident on this GPU?  nvcc -arch=sm_100a -o max_clusters max_clusters.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void stub(float* p) { extern __shared__ float s[]; if (p) p[0] = s[0]; }

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int smem = 230000;
  cudaFuncSetAttribute(stub, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(stub, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(prop.multiProcessorCount / cs * cs);
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, stub, &cfg);
    printf("cluster size %2d: max active clusters %d (%d SMs)  %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
