# NetworkSolversB200.jl -- Julia host shim binding libnsb200.so (include/nsb200.h) behind the problem-type seam of
# NetworkSolvers.jl.  The sweep driver, iterators, region plans and kwarg packs of the reference stay untouched
# (src/iterators.jl:81-104 dispatches extracter / updater / inserter / region_plan on the problem type); the methods below
# forward the three hooks to the device library, where the state, the projected operator and the local tensor live.
#
# STATUS: written against the C ABI; Julia is not available in the build image (no network), so this file has not been
# executed here.  Every ccall signature below is the one the Python host (networksolvers_b200/_lib.py) binds and the GPU test
# suite exercises; tests/test_julia_shim.py checks that each symbol and struct layout used here exists in the header.
#
#   using NetworkSolversB200
#   E, psi = NetworkSolversB200.dmrg(H, psi0; nsweeps=5, nsites=2, inserter_kwargs=(; trunc=(; cutoff=1e-12, maxdim=100)))
#   NetworkSolversB200.dmrg(H, psi0; devices=0:7, ...)        # one Julia process, eight GPUs (nsb_multi_*)
module NetworkSolversB200

import NetworkSolvers as ns
import ITensors as it
import ITensorNetworks as itn
import Graphs
using ConstructionBase: setproperties

const lib = get(ENV, "NSB200_LIB", "libnsb200.so")

# ---- PODs of include/nsb200.h (field order and types are the C struct's) ------------------------------------------
struct NsbTrunc;  cutoff::Cdouble; mindim::Int64; maxdim::Int64; end
struct NsbExpand; algorithm::Int32; north_pass::Int32; expansion_factor::Cdouble; max_expand::Int64; end
struct NsbKrylov; krylovdim::Int32; maxiter::Int32; tol::Cdouble; which::Int32; eager::Int32; rk_order::Int32; reserved::Int32; end
struct NsbExtractInfo; expanded::Int32; env_builds::Int32; qr_steps::Int32; local_rank::Int32; local_numel::Int64; end
struct NsbSolveInfo;   nmatvec::Int32; krylovdim::Int32; converged::Int32; reserved::Int32; residual::Cdouble; end
struct NsbInsertInfo;  newdim::Int64; truncerr::Cdouble; decomp::Int32; jacobi_sweeps::Int32; end

const NSB_F64, NSB_C128 = Int32(0), Int32(1)
const NSB_SITE, NSB_SITE_OUT = Int32(-1), Int32(-2)
const NSB_SOLVER_RK, NSB_SOLVER_KRYLOV = Int32(0), Int32(1)
const NSB_EXPAND_DENSITYMATRIX, NSB_EXPAND_ORTHO = Int32(1), Int32(2)
clampi64(x) = x >= typemax(Int64) ? typemax(Int64) : Int64(x)

# ---- device network ----------------------------------------------------------------------------------------------
# One handle type for both ways of running: a single context (ctx, net) or a multi-device group (multi != C_NULL) whose hooks
# fan out inside the library (nsb_multi_*); net is then device 0's replica, used for the non-collective queries.
mutable struct DeviceNet
  ctx::Ptr{Cvoid}
  net::Ptr{Cvoid}
  multi::Ptr{Cvoid}
  graph                      # the NamedGraph of the state (vertex names as in the ITensorNetwork)
  verts::Vector{Any}         # vertex id (0-based position) -> vertex name
  vid::Dict{Any,Int32}
  siteinds::Dict{Any,Any}    # vertex -> site Index (kept on the host: downloads are rebuilt on the same indices)
  linkinds::Dict{Any,Any}    # (u, v) -> link Index as last seen on the host (dimension changes are re-made on download)
  eltype::DataType
  function DeviceNet(ctx, net, multi, graph, verts, vid, sinds, linds, elt)
    d = new(ctx, net, multi, graph, verts, vid, sinds, linds, elt)
    finalizer(close!, d)
    return d
  end
end

function check(d::DeviceNet, rc)
  rc == 0 && return nothing
  msg = d.multi != C_NULL ? unsafe_string(ccall((:nsb_multi_last_error, lib), Cstring, (Ptr{Cvoid},), d.multi)) :
                            unsafe_string(ccall((:nsb_last_error, lib), Cstring, (Ptr{Cvoid},), d.ctx))
  error("libnsb200 error $rc: $msg")
end

function close!(d::DeviceNet)
  if d.multi != C_NULL
    ccall((:nsb_multi_destroy, lib), Cint, (Ptr{Cvoid},), d.multi); d.multi = C_NULL
  elseif d.net != C_NULL
    ccall((:nsb_network_destroy, lib), Cint, (Ptr{Cvoid},), d.net)
    ccall((:nsb_ctx_destroy, lib), Cint, (Ptr{Cvoid},), d.ctx)
  end
  d.net = C_NULL; d.ctx = C_NULL
  return nothing
end

graph(d::DeviceNet) = d.graph
vertex_ids(d::DeviceNet, region) = Int32[d.vid[v] for v in region]

# Leg encoding of a tensor at vertex v: the tensor is handed over in the order permute_indices gives it
# (src/permute_indices.jl:11-15: first link, site indices, other links) -- the library keeps exactly that order.
#   site index            -> (v, NSB_SITE);   primed site index (operators) -> (v, NSB_SITE_OUT)
#   link to neighbour n   -> (v, n)
function ordered_inds(tn, v; operator=false)
  nbrs = collect(Graphs.neighbors(tn, v))
  links = [only(it.commoninds(tn[v], tn[n])) for n in nbrs]
  sites = [i for i in it.inds(tn[v]) if !(i in links)]
  operator && (sites = sort(sites; by=it.plev))            # (s, s')
  order = isempty(links) ? sites : vcat(links[1:1], sites, links[2:end])
  return order, nbrs, links
end

function encode_legs(d::DeviceNet, tn, v; operator=false)
  order, nbrs, links = ordered_inds(tn, v; operator)
  legs = Int32[]
  for i in order
    k = findfirst(==(i), links)
    if isnothing(k)
      append!(legs, (d.vid[v], it.plev(i) == 0 ? NSB_SITE : NSB_SITE_OUT))
    else
      append!(legs, (d.vid[v], d.vid[nbrs[k]]))
    end
  end
  return order, legs
end

function upload!(d::DeviceNet, tn, v; operator=false)
  order, legs = encode_legs(d, tn, v; operator)
  A = Array{d.eltype}(it.array(it.dense(tn[v]), order...))     # dense column-major copy in that index order (QN tensors: dense image)
  dims = collect(Int64, size(A))
  # (the symbol of a ccall must be a literal: one call per entry point)
  m, n, vv, r = d.multi, d.net, d.vid[v], Int32(ndims(A))
  rc = if d.multi != C_NULL && operator
    ccall((:nsb_multi_mpo_upload, lib), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Ptr{Int64}, Ptr{Cvoid}), m, vv, r, legs, dims, A)
  elseif d.multi != C_NULL
    ccall((:nsb_multi_site_upload, lib), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Ptr{Int64}, Ptr{Cvoid}), m, vv, r, legs, dims, A)
  elseif operator
    ccall((:nsb_mpo_upload, lib), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Ptr{Int64}, Ptr{Cvoid}), n, vv, r, legs, dims, A)
  else
    ccall((:nsb_site_upload, lib), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Ptr{Int64}, Ptr{Cvoid}), n, vv, r, legs, dims, A)
  end
  check(d, rc)
end

# abelian quantum numbers: one row of integer charges per basis state of every site and link index (conserve_qns = true,
# examples/dmrg.jl:10).  Link charges are those of the subtree on the first vertex's side, as nsb_qn_set_link documents.
qn_names(i::it.Index) = unique(vcat([[String(it.name(q)) for q in it.qn(s).data if it.name(q) != it.SmallString("")] for s in it.space(i)]...))
function charges(i::it.Index, names; flip=false)
  rows = Vector{Int32}[]
  for (q, dim) in it.space(i)
    c = Int32[(flip ? -1 : 1) * Int(it.dir(i)) * it.val(q, n) for n in names]
    for _ in 1:dim; push!(rows, c); end
  end
  return permutedims(reduce(hcat, rows))            # dim x nq
end
function upload_qns!(d::DeviceNet, psi, total)
  s1 = d.siteinds[first(d.verts)]
  names = qn_names(s1)
  tot = Int32[it.val(total, n) for n in names]
  check(d, ccall((:nsb_qn_enable, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}), d.net, length(names), tot))
  for v in d.verts
    c = permutedims(charges(d.siteinds[v], names))   # state-major rows
    check(d, ccall((:nsb_qn_set_site, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}), d.net, d.vid[v], c))
  end
  for e in Graphs.edges(psi)
    u, v = Graphs.src(e), Graphs.dst(e)
    l = only(it.commoninds(psi[u], psi[v]))
    c = permutedims(charges(l, names; flip=(it.dir(it.inds(psi[u])[findfirst(==(l), it.inds(psi[u]))]) == it.In)))
    check(d, ccall((:nsb_qn_set_link, lib), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}), d.net, d.vid[u], d.vid[v], c))
  end
end

"""
    DeviceNet(H, psi; devices=[0])

Counterpart of `permute_indices(init_state)`, `permute_indices(H)`, `itn.ProjTTN(H)` in `src/eigsolve.jl:69-74` /
`src/applyexp.jl:84-89`: uploads the state and the operator, records the orthogonality region; the projected operator's
environments are built lazily on the device by the first `extracter`.
"""
function DeviceNet(H, psi; devices=[0], eltype=nothing)
  g = itn.underlying_graph(psi)                 # the NamedGraph (examples/quench_evolution.jl:33 uses the same call)
  verts = collect(Graphs.vertices(psi))
  vid = Dict{Any,Int32}(v => Int32(i - 1) for (i, v) in enumerate(verts))
  elt = isnothing(eltype) ? promote_type(it.scalartype(psi), it.scalartype(H)) : eltype
  elt = elt <: Complex ? ComplexF64 : Float64
  dtype = elt <: Complex ? NSB_C128 : NSB_F64
  edges = Int32[]
  for e in Graphs.edges(psi); append!(edges, (vid[Graphs.src(e)], vid[Graphs.dst(e)])); end
  si = itn.siteinds(psi)                          # as src/permute_indices.jl:5 reads them
  sinds = Dict{Any,Any}(v => only(si[v]) for v in verts)
  sdims = Int64[it.dim(sinds[v]) for v in verts]
  linds = Dict{Any,Any}()
  for e in Graphs.edges(psi)
    u, v = Graphs.src(e), Graphs.dst(e)
    linds[(u, v)] = linds[(v, u)] = only(it.commoninds(psi[u], psi[v]))
  end
  devs = collect(Int32, devices)
  ctx = Ref{Ptr{Cvoid}}(C_NULL); net = Ref{Ptr{Cvoid}}(C_NULL); multi = Ref{Ptr{Cvoid}}(C_NULL)
  if length(devs) == 1
    rc = ccall((:nsb_ctx_create, lib), Cint, (Cint, Ptr{Ptr{Cvoid}}), devs[1], ctx)
    rc == 0 || error("nsb_ctx_create failed ($rc): " * unsafe_string(ccall((:nsb_last_error, lib), Cstring, (Ptr{Cvoid},), C_NULL)))
    rc = ccall((:nsb_network_create, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Int64}, Int32, Ptr{Ptr{Cvoid}}),
               ctx[], length(verts), edges, length(edges) ÷ 2, sdims, dtype, net)
    rc == 0 || error("nsb_network_create failed ($rc)")
  else
    rc = ccall((:nsb_multi_create, lib), Cint, (Ptr{Int32}, Int32, Ptr{Ptr{Cvoid}}), devs, length(devs), multi)
    rc == 0 || error("nsb_multi_create failed ($rc)")
    rc = ccall((:nsb_multi_network_create, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Int64}, Int32),
               multi[], length(verts), edges, length(edges) ÷ 2, sdims, dtype)
    rc == 0 || error("nsb_multi_network_create failed ($rc)")
    ccall((:nsb_multi_ctx, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}), multi[], 0, ctx)
    ccall((:nsb_multi_net, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}), multi[], 0, net)
  end
  d = DeviceNet(ctx[], net[], multi[], g, verts, vid, sinds, linds, elt)
  Hp, psip = ns.permute_indices(H), ns.permute_indices(psi)
  for v in verts
    upload!(d, Hp, v; operator=true)
    upload!(d, psip, v)
  end
  region = vertex_ids(d, collect(itn.ortho_region(psi)))
  if d.multi != C_NULL
    check(d, ccall((:nsb_multi_set_ortho_region, lib), Cint, (Ptr{Cvoid}, Ptr{Int32}, Int32), d.multi, region, length(region)))
  else
    check(d, ccall((:nsb_set_ortho_region, lib), Cint, (Ptr{Cvoid}, Ptr{Int32}, Int32), d.net, region, length(region)))
  end
  if it.hasqns(psi[first(verts)])
    d.multi == C_NULL || error("QN conservation with devices > 1: upload the charges to every replica (nsb_multi_net) -- not wired in this shim")
    upload_qns!(d, psip, reduce(+, [it.flux(psip[v]) for v in verts]))      # total charge = sum of the tensors' fluxes
  end
  return d
end

# ---- state download (callbacks such as examples/quench_evolution.jl:44-47 read problem.state) -----------------------
function linkdim(d::DeviceNet, u, v)
  r = Ref{Int64}()
  check(d, ccall((:nsb_linkdim, lib), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Int64}), d.net, d.vid[u], d.vid[v], r))
  return r[]
end
function maxlinkdim(d::DeviceNet)
  r = Ref{Int64}()
  check(d, ccall((:nsb_maxlinkdim, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}), d.net, r))
  return r[]
end
itn.maxlinkdim(d::DeviceNet) = maxlinkdim(d)

"""
    download_state(d) -> TreeTensorNetwork

Every site tensor comes back in the library's canonical order (first link, site, other links); link indices whose dimension
changed on the device are re-made (same tags), all others are the host's original `Index` objects.
"""
function download_state(d::DeviceNet)
  for e in Graphs.edges(d.graph)
    u, v = Graphs.src(e), Graphs.dst(e)
    n = linkdim(d, u, v)
    if it.dim(d.linkinds[(u, v)]) != n
      d.linkinds[(u, v)] = d.linkinds[(v, u)] = it.Index(n; tags=it.tags(d.linkinds[(u, v)]))
    end
  end
  tensors = Dict{Any,it.ITensor}()
  for v in d.verts
    rank = Ref{Int32}(); legs = zeros(Int32, 32); dims = zeros(Int64, 16)
    check(d, ccall((:nsb_site_info, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}), d.net, d.vid[v], rank, legs, dims))
    r = Int(rank[])
    A = Array{d.eltype}(undef, dims[1:r]...)
    check(d, ccall((:nsb_site_download, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Cvoid}), d.net, d.vid[v], A))
    inds = map(1:r) do k
      b = legs[2k]
      b == NSB_SITE ? d.siteinds[v] : d.linkinds[(v, d.verts[b + 1])]
    end
    tensors[v] = it.ITensor(A, inds...)
  end
  psi = itn.TreeTensorNetwork(itn.ITensorNetwork(tensors))          # same graph as d.graph
  region = Ref{Int32}(); vs = zeros(Int32, length(d.verts))
  check(d, ccall((:nsb_get_ortho_region, lib), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), d.net, vs, region))
  return itn.set_ortho_region(psi, [d.verts[vs[k] + 1] for k in 1:region[]])
end

# ---- problem types -------------------------------------------------------------------------------------------------
struct LocalState; net::DeviceNet; end               # token: the local tensor stays on the device between the hooks
function Base.Array(l::LocalState)
  rank = Ref{Int32}(); legs = zeros(Int32, 32); dims = zeros(Int64, 16)
  check(l.net, ccall((:nsb_local_info, lib), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}), l.net.net, rank, legs, dims))
  A = Array{l.net.eltype}(undef, dims[1:rank[]]...)
  if l.net.multi != C_NULL
    check(l.net, ccall((:nsb_multi_local_download, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), l.net.multi, A))
  else
    check(l.net, ccall((:nsb_local_download, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), l.net.net, A))
  end
  return A
end

@kwdef struct B200EigsolveProblem
  net::DeviceNet
  eigenvalue::Number = Inf
end
@kwdef struct B200ApplyExpProblem
  net::DeviceNet
  current_time::Number = 0.0
end
@kwdef struct B200FittingProblem
  net::DeviceNet                 # state = ket being fitted, operator = A (identity network for itn.truncate), target uploaded
  overlap::Float64 = 0.0
end
const B200Problem = Union{B200EigsolveProblem,B200ApplyExpProblem,B200FittingProblem}

ns.eigenvalue(E::B200EigsolveProblem) = E.eigenvalue
ns.current_time(T::B200ApplyExpProblem) = T.current_time
ns.state(P::B200Problem) = download_state(P.net)            # lazy: only when a callback or the caller asks
ns.operator(P::B200Problem) = P.net
ns.region_plan(T::B200ApplyExpProblem; nsites, time_step, sweep_kwargs...) =
  ns.tdvp_regions(graph(T.net), time_step; nsites, sweep_kwargs...)                     # src/applyexp.jl:14-16
ns.region_plan(P::Union{B200EigsolveProblem,B200FittingProblem}; nsites, sweep_kwargs...) =
  ns.euler_sweep(graph(P.net); nsites, sweep_kwargs...)                                  # src/iterators.jl:102-104

# The reference's printers call itn.maxlinkdim(state(E)) (src/eigsolve.jl:39, src/applyexp.jl:56), which through `ns.state` above
# would download the whole state once per sweep.  The printers below answer from the device; they are separate functions (the
# reference's own methods are left alone) and the default `sweep_printer` of this module's entry points.
function eigsolve_sweep_printer(region_iterator; outputlevel, sweep, nsweeps, kws...)
  outputlevel >= 1 || return nothing
  E = ns.problem(region_iterator)
  println("After sweep $sweep/$nsweeps eigenvalue=$(E.eigenvalue) maxlinkdim=$(maxlinkdim(E.net))")
  flush(stdout)
end
function applyexp_sweep_printer(region_iterator; outputlevel, sweep, nsweeps, process_time=identity, kws...)
  outputlevel >= 1 || return nothing
  T = ns.problem(region_iterator)
  println("  Current time = $(process_time(T.current_time)), maxlinkdim=$(maxlinkdim(T.net))")
  flush(stdout)
end

# ---- hooks ------------------------------------------------------------------------------------------------------------
function ns.extracter(P::B200Problem, region_iterator; sweep, trunc=(;), subspace_algorithm=nothing, north_pass=1,
                      expansion_factor=ns.default_expansion_factor(), max_expand=ns.default_max_expand(), kws...)
  t = ns.truncation_parameters(sweep; trunc...)
  region = vertex_ids(P.net, ns.current_region(region_iterator))
  tr = Ref(NsbTrunc(t.cutoff, t.mindim, clampi64(t.maxdim)))
  alg = isnothing(subspace_algorithm) || P isa B200FittingProblem ? Int32(0) :       # src/fitting.jl:38: no expansion when fitting
        subspace_algorithm == "densitymatrix" ? NSB_EXPAND_DENSITYMATRIX :
        (subspace_algorithm == "ortho" && P isa B200EigsolveProblem) ? NSB_EXPAND_ORTHO :
        error("Subspace expansion (subspace_expand!) not defined for requested combination of subspace_algorithm and problem types")
  ex = Ref(NsbExpand(alg, north_pass, expansion_factor, clampi64(max_expand)))
  info = Ref{NsbExtractInfo}()
  d = P.net
  if d.multi != C_NULL
    check(d, ccall((:nsb_multi_extract, lib), Cint, (Ptr{Cvoid}, Ptr{Int32}, Int32, Ptr{NsbTrunc}, Ptr{NsbExpand}, Ptr{NsbExtractInfo}),
                   d.multi, region, length(region), tr, alg == 0 ? C_NULL : ex, info))
    check(d, ccall((:nsb_multi_set_shard, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}), d.multi, 1, C_NULL))   # shard this position
  else
    check(d, ccall((:nsb_extract, lib), Cint, (Ptr{Cvoid}, Ptr{Int32}, Int32, Ptr{NsbTrunc}, Ptr{NsbExpand}, Ptr{NsbExtractInfo}),
                   d.net, region, length(region), tr, alg == 0 ? C_NULL : ex, info))
  end
  return P, LocalState(d)
end

function ns.updater(E::B200EigsolveProblem, local_state, region_iterator; outputlevel, solver=ns.eigsolve_solver,
                    which_eigval=:SR, tol=1e-14, krylovdim=3, maxiter=1, eager=false, kws...)
  solver === ns.eigsolve_solver || error("B200EigsolveProblem runs eigsolve_solver on the device; other solvers need the CPU problem type")
  kp = Ref(NsbKrylov(krylovdim, maxiter, tol, which_eigval == :SR ? 0 : 1, eager, 4, 0))
  val = Ref{Cdouble}(); info = Ref{NsbSolveInfo}()
  d = E.net
  if d.multi != C_NULL
    check(d, ccall((:nsb_multi_update_eigsolve, lib), Cint, (Ptr{Cvoid}, Ptr{NsbKrylov}, Ptr{Cdouble}, Ptr{NsbSolveInfo}), d.multi, kp, val, info))
  else
    check(d, ccall((:nsb_update_eigsolve, lib), Cint, (Ptr{Cvoid}, Ptr{NsbKrylov}, Ptr{Cdouble}, Ptr{NsbSolveInfo}), d.net, kp, val, info))
  end
  outputlevel >= 2 && println("  Region $(ns.current_region(region_iterator)): energy = $(val[])")
  return setproperties(E; eigenvalue=val[]), local_state
end

# vertex the 1-site TDVP step splits toward: the first hop of the path from the current region to the next one
# (src/applyexp.jl:30-36); -1 when nsites != 1 or there is no next region
function next_hop(d::DeviceNet, region_iterator, nsites)
  nsites == 1 || return Int32(-1)
  curr, nxt = ns.current_region(region_iterator), ns.next_region(region_iterator)
  (isnothing(nxt) || nxt == curr) && return Int32(-1)
  next_edge = first(itn.edge_sequence_between_regions(d.graph, curr, nxt))    # the reference's own call (src/applyexp.jl:34)
  return d.vid[Graphs.dst(next_edge)]
end

function ns.updater(T::B200ApplyExpProblem, local_state, region_iterator; nsites, time_step, solver=ns.runge_kutta_solver,
                    outputlevel, order=4, krylovdim=30, maxiter=100, tol=1e-12, eager=true, kws...)
  code = solver === ns.runge_kutta_solver ? NSB_SOLVER_RK :
         solver === ns.exponentiate_solver ? NSB_SOLVER_KRYLOV : error("ApplyExpProblem on the device needs runge_kutta_solver or exponentiate_solver")
  code == NSB_SOLVER_RK && !(order in (2, 4)) && error("For runge_kutta_solver, must specify `order` keyword")
  kp = Ref(NsbKrylov(krylovdim, maxiter, tol, 0, eager, order, 0)); info = Ref{NsbSolveInfo}()
  d = T.net
  nxt = next_hop(d, region_iterator, nsites)
  if d.multi != C_NULL
    check(d, ccall((:nsb_multi_update_exp, lib), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Int32, Ptr{NsbKrylov}, Int32, Int32, Ptr{NsbSolveInfo}),
                   d.multi, real(time_step), imag(time_step), code, kp, nsites, nxt, info))
  else
    check(d, ccall((:nsb_update_exp, lib), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Int32, Ptr{NsbKrylov}, Int32, Int32, Ptr{NsbSolveInfo}),
                   d.net, real(time_step), imag(time_step), code, kp, nsites, nxt, info))
  end
  return setproperties(T; current_time=T.current_time + time_step), local_state
end

function ns.updater(F::B200FittingProblem, local_state, region_iterator; outputlevel, kws...)     # src/fitting.jl:42-49
  ov = Ref{Cdouble}()
  check(F.net, ccall((:nsb_update_fit, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), F.net.net, ov))
  outputlevel >= 2 && println("  Region $(ns.current_region(region_iterator)): squared overlap = $(ov[])")
  return setproperties(F; overlap=ov[]), local_state
end

function ns.inserter(P::B200Problem, local_tensor, region_iterator; normalize=false, set_orthogonal_region=true, sweep,
                     trunc=(;), kws...)
  t = ns.truncation_parameters(sweep; trunc...)
  n = length(ns.current_region(region_iterator))
  n in (1, 2) || error("Region of length $n not currently supported")                   # src/inserter.jl:26
  tr = Ref(NsbTrunc(t.cutoff, t.mindim, clampi64(t.maxdim))); info = Ref{NsbInsertInfo}()
  d = P.net
  if d.multi != C_NULL
    check(d, ccall((:nsb_multi_insert, lib), Cint, (Ptr{Cvoid}, Ptr{NsbTrunc}, Int32, Int32, Ptr{NsbInsertInfo}),
                   d.multi, tr, normalize, set_orthogonal_region, info))
  else
    check(d, ccall((:nsb_insert, lib), Cint, (Ptr{Cvoid}, Ptr{NsbTrunc}, Int32, Int32, Ptr{NsbInsertInfo}),
                   d.net, tr, normalize, set_orthogonal_region, info))
  end
  return P
end

# ---- entry points with the reference's signatures ---------------------------------------------------------------------
function eigsolve(H, init_state; devices=[0], sweep_printer=eigsolve_sweep_printer, kws...)   # src/eigsolve.jl:69-74
  # (eigenvalue, state) as the reference returns; the state is downloaded once, at the end
  return ns.eigsolve(B200EigsolveProblem(; net=DeviceNet(H, init_state; devices)); sweep_printer, kws...)
end
dmrg(args...; kws...) = eigsolve(args...; kws...)

function applyexp(H, init_state, exponents; devices=[0], sweep_printer=applyexp_sweep_printer, kws...)   # src/applyexp.jl:84-89
  prob = B200ApplyExpProblem(; net=DeviceNet(H, init_state; devices, eltype=ComplexF64))
  return ns.applyexp(prob, exponents; sweep_printer, kws...)
end
function tdvp(H, init_state, time_points; process_time=ns.process_real_times,
              sweep_printer=(a...; k...) -> applyexp_sweep_printer(a...; process_time, k...), kws...)   # src/applyexp.jl:93-103
  return applyexp(H, init_state, [-im * t for t in time_points]; sweep_printer, kws...)
end

function fit_tensornetwork(target, operator, init_state; nsweeps=25, nsites=1, outputlevel=0, normalize=true,
                           extracter_kwargs=(;), updater_kwargs=(;), inserter_kwargs=(;), kws...)   # src/fitting.jl:55-84
  d = DeviceNet(operator, init_state)
  tp = ns.permute_indices(target)
  for v in d.verts
    order, legs = encode_legs(d, tp, v)
    A = Array{d.eltype}(it.array(tp[v], order...)); dims = collect(Int64, size(A))
    check(d, ccall((:nsb_fit_target_upload, lib), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Ptr{Int64}, Ptr{Cvoid}),
                   d.net, d.vid[v], ndims(A), legs, dims, A))
  end
  ik = (; inserter_kwargs..., normalize, set_orthogonal_region=false)
  sweep_iter = ns.sweep_iterator(B200FittingProblem(; net=d), nsweeps; nsites, outputlevel, extracter_kwargs, updater_kwargs, inserter_kwargs=ik)
  return ns.state(ns.sweep_solve(sweep_iter; outputlevel, kws...))
end

end # module
