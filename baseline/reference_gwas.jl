# Times the UNMODIFIED reference (GenomicBreedingModels.jl, /root/reference/src/gwas.jl:206-259, :329-399) on the
# host cores, when a Julia runtime with the package and its dependencies exists on the box.  bench.py --impl reference
# probes for `julia` (PATH, baseline/_ref/) and runs this script; without one it times the oracle's C/OpenMP port
# instead and says so.  NOT EXECUTED IN THE BUILD IMAGE (no Julia there).
#
#   julia --threads=auto,1 baseline/reference_gwas.jl <n> <l> <model: ols|lmm> <reps>
#
# Prints one line: "reference_gwas n l model seconds_per_call n_threads blas_threads l_kept"
using GenomicBreedingCore, GenomicBreedingModels, LinearAlgebra
n, l, model, reps = parse(Int, ARGS[1]), parse(Int, ARGS[2]), ARGS[3], parse(Int, ARGS[4])
genomes = GenomicBreedingCore.simulategenomes(n = n, l = l, verbose = false)       # as the doctests, gwas.jl:41
ploidy = 4
genomes.allele_frequencies = round.(genomes.allele_frequencies .* ploidy) ./ ploidy  # gwas.jl:43-45
proportion_of_variance = zeros(9, 1); proportion_of_variance[1, 1] = 0.5             # gwas.jl:47
trials, _ = GenomicBreedingCore.simulatetrials(genomes = genomes, n_years = 1, n_seasons = 1, n_harvests = 1, n_sites = 1,
    n_replications = 1, f_add_dom_epi = [0.05 0.00 0.00;], proportion_of_variance = proportion_of_variance, verbose = false)
phenomes = extractphenomes(trials)
f = model == "lmm" ? gwaslmm : gwasols
fit = f(genomes = genomes, phenomes = phenomes, GRM_type = "simple")               # JIT warm-up call
t = @elapsed for _ in 1:reps
    global fit = f(genomes = genomes, phenomes = phenomes, GRM_type = "simple")
end
println("reference_gwas ", n, " ", l, " ", model, " ", t / reps, " ", Threads.nthreads(), " ", BLAS.get_num_threads(), " ", length(fit.b_hat))
