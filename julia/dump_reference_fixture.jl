# dump_reference_fixture.jl -- pins the CUDA path / the oracle on the UNMODIFIED reference.
#
# Runs RankCompV3.identify_degs (src/RankCompV3.jl:339-438, unexported, called exactly as reoa() does at src:652-662)
# on the tie-free inputs written by scripts/make_tiefree_inputs.py and dumps its complete return value.  On tie-free
# input is_greater never reaches rand(Bool) (src:72-73), so the reference is deterministic and its output can be
# compared bit for bit (tables, calls) / to 1e-12 (p-values) with libreo_cuda.so:
#     tests/test_gpu_parity.py::test_reference_julia_fixtures   consumes tests/golden/julia_out/<case>.tsv when present.
#
# Usage (a machine with Julia >= 1.7 and the reference checked out; none of this can run in the build image):
#     julia --project=/path/to/RankCompV3.jl julia/dump_reference_fixture.jl tests/golden/julia_in tests/golden/julia_out
#
# Output, one file per case: tab separated, no header, r rows x (1 + 16K) columns = the `res` matrix of src:394/430
# (gene name; per level k: pval padj n11 n12 n13 n21 n22 n23 n31 n32 n33 d1 d2 se z1 up_down), floats printed with
# 17 significant digits so that they round-trip.
using DelimitedFiles
using Printf
import RankCompV3

indir, outdir = ARGS[1], ARGS[2]
mkpath(outdir)
for f in sort(readdir(indir))
    endswith(f, "_expr.tsv") || continue
    case = replace(f, "_expr.tsv" => "")
    expr = readdlm(joinpath(indir, f), '\t', Any; header = true)[1]
    genes = String.(expr[:, 1])
    data = Matrix{Int64}(expr[:, 2:end])                       # Matrix(df_expr) of src:652: r x c Int64
    meta = readdlm(joinpath(indir, case * "_meta.tsv"), '\t', String; header = true)[1]
    group = meta[:, 2]                                         # meta_group.Group, by position (src:614-615)
    refm = readdlm(joinpath(indir, case * "_ref.tsv"), '\t', Any; header = true)[1]
    ref_gene = BitVector(Int.(refm[:, 2]) .!= 0)               # ref_gene_vec of src:651
    par = readdlm(joinpath(indir, case * "_par.tsv"), '\t', Any; header = true)[1]
    pval_reo, pval_deg, padj_deg = Float64(par[1, 1]), Float64(par[1, 2]), Float64(par[1, 3])
    n_iter, n_conv = Int(par[1, 4]), Int(par[1, 5])
    res = RankCompV3.identify_degs(data, group, genes, pval_reo, pval_deg, padj_deg, ref_gene, n_iter, n_conv)
    open(joinpath(outdir, case * ".tsv"), "w") do io
        for i in 1:size(res, 1)
            cells = String[]
            for v in res[i, :]
                push!(cells, v isa AbstractFloat ? @sprintf("%.17g", v) : string(v))
            end
            println(io, join(cells, '\t'))
        end
    end
    println("wrote ", joinpath(outdir, case * ".tsv"), "  ", size(res))
end
