import json, sys, glob
for f in sorted(glob.glob(sys.argv[1])):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    k = d["kernels"]
    print("%-40s step %.3f ms | pairs %.4f lin %.4f backsub %.4f chol %.4f tri %.4f fin %.4f | parity %s" % (
        f.split("/")[-1], d["ms_per_step"], k["k_schur_pairs"]["ms_avg"], k["k_lin_points"]["ms_avg"], k["k_backsub"]["ms_avg"],
        k["chol_graph"]["ms_avg"], k["k_tri_solve"]["ms_avg"], k["k_S_finalize"]["ms_avg"], d["parity_vs_1gpu"]["ok"] if d.get("parity_vs_1gpu") else None))
