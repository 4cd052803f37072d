"""gbm_b200: host-side mirror of the GenomicBreedingModels.jl GWAS / GRM hot path over
libgbm_b200.so (hand-written CUDA for sm_100a).  See ../../include/gbm_b200.h for the C ABI
and ../julia/GenomicBreedingModelsB200.jl for the Julia shim over the same symbols."""
from . import _lib
from ._lib import ArgumentError, CudaError, ErrorException, build, init, last_timing, load
from .core import DeviceMatrix, LmmPlan, ScanPlan, gemm_tn, grm_finalize, kstd_pc1, kstd_pc1_device, measure_copy_bandwidth, neglog10_sf, pack_host, scan_host
from .gwas import extractxyetc, grmploidyaware, grmsimple, gwaslmm, gwasols, gwasprep, gwasreml
from .structs import GRM, Fit, Genomes, Phenomes
from . import transform
from .transform import epistasisfeatures, transform1, transform2

__all__ = [
    "ArgumentError", "CudaError", "ErrorException", "build", "init", "load", "last_timing", "DeviceMatrix", "LmmPlan", "ScanPlan", "gemm_tn",
    "grm_finalize", "kstd_pc1", "kstd_pc1_device", "measure_copy_bandwidth", "neglog10_sf", "pack_host", "scan_host",
    "extractxyetc", "grmploidyaware", "grmsimple", "gwaslmm", "gwasols", "gwasprep", "gwasreml", "GRM", "Fit", "Genomes",
    "Phenomes", "transform", "transform1", "transform2", "epistasisfeatures",
]
