#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference (manodeep/Corrfunc v2.5.3)
# pair-counting libraries from the sources where they lie under /root/reference into
# oracle/_ref/libcorrfunc_ref_{v4,v3}.so.  Nothing from the reference is copied into the repo:
# the precision templates (*.c.src / *.h.src) are expanded with the same `sed` rule the
# reference's rules.mk uses (rules.mk:23-49) into a throw-away temp dir that is removed on exit.
#
# We do NOT run the reference's own build system; this is a direct gcc recipe with the
# reference's own flags (common.mk:175-176,243,257,333).  Two variants are built because
# /root/reference does not exist on the GPU box and the box's CPU may differ from this one:
#   _v4 : -march=x86-64-v4  (AVX-512F kernels compiled in; used when the host has avx512f)
#   _v3 : -march=x86-64-v3  (AVX2/FMA host; the reference then dispatches to its AVX kernels)
# The reference picks kernels at run time from options->instruction_set, capped by cpuid
# (theory/DD/countpairs_impl.c.src:40-133).
set -euo pipefail
REF=${CORRFUNC_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/theory/DD" ]; then
  echo "build_ref: $REF not present; keeping any prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d /tmp/corrfunc_ref_build.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT

expand() { # expand <src template> <dst dir>  -> writes <name>_float.<ext> and <name>_double.<ext>
  local src="$1" dst="$2" base ext
  base="$(basename "$src" .src)"; ext="${base##*.}"; base="${base%.*}"
  { echo "#ifdef DOUBLE_PREC"; echo "#undef DOUBLE_PREC"; echo "#endif";
    sed -e "/DOUBLE_PREC/!s/DOUBLE/float/g" "$src"; } > "$dst/${base}_float.${ext}"
  { echo "#ifndef DOUBLE_PREC"; echo "#define DOUBLE_PREC"; echo "#endif";
    sed -e "/DOUBLE_PREC/!s/DOUBLE/double/g" "$src"; } > "$dst/${base}_double.${ext}"
}

mkdir -p "$TMP/gen"
for d in utils theory/vpf theory/DD theory/DDrppi theory/DDsmu theory/wp theory/xi mocks/DDtheta_mocks mocks/DDrppi_mocks mocks/DDsmu_mocks mocks/vpf_mocks; do
  for f in "$REF/$d"/*.src; do expand "$f" "$TMP/gen"; done
done

INCL="-I$TMP/gen -I$REF/utils -I$REF/io -I$REF/theory/vpf -I$REF/theory/DD -I$REF/theory/DDrppi -I$REF/theory/DDsmu -I$REF/theory/wp -I$REF/theory/xi -I$REF/mocks/DDtheta_mocks -I$REF/mocks/DDrppi_mocks -I$REF/mocks/DDsmu_mocks -I$REF/mocks/vpf_mocks -I$HERE/gsl_shim"
COMMON="-std=c99 -m64 -O3 -fPIC -D_POSIX_SOURCE=200809L -D_GNU_SOURCE -DVERSION=\"2.5.3\" -DUSE_OMP -fopenmp \
 -funroll-loops -fno-strict-aliasing -ftree-vectorize -DPERIODIC -DENABLE_MIN_SEP_OPT -DCOPY_PARTICLES -DOUTPUT_RPAVG \
 -DLINK_IN_DEC -DLINK_IN_RA -DDOUBLE_PREC -w"

SRCS=(
  "$REF/theory/DD/countpairs.c" "$TMP/gen/countpairs_impl_float.c" "$TMP/gen/countpairs_impl_double.c"
  "$REF/theory/DDrppi/countpairs_rp_pi.c" "$TMP/gen/countpairs_rp_pi_impl_float.c" "$TMP/gen/countpairs_rp_pi_impl_double.c"
  "$REF/theory/DDsmu/countpairs_s_mu.c" "$TMP/gen/countpairs_s_mu_impl_float.c" "$TMP/gen/countpairs_s_mu_impl_double.c"
  "$REF/theory/wp/countpairs_wp.c" "$TMP/gen/countpairs_wp_impl_float.c" "$TMP/gen/countpairs_wp_impl_double.c"
  "$REF/theory/xi/countpairs_xi.c" "$TMP/gen/countpairs_xi_impl_float.c" "$TMP/gen/countpairs_xi_impl_double.c"
  "$REF/mocks/DDtheta_mocks/countpairs_theta_mocks.c" "$TMP/gen/countpairs_theta_mocks_impl_float.c" "$TMP/gen/countpairs_theta_mocks_impl_double.c"
  # SURVEY 8(f) rank 1.  GSL is absent: gsl_shim/ stands in for <gsl/gsl_interp.h> (linear interpolation, restated from
  # GSL's source; only reached with is_comoving_dist == 0) and for the unused <gsl/gsl_integration.h> include of
  # set_cosmo_dist.c; the reference sources themselves are compiled unmodified.
  "$REF/mocks/DDrppi_mocks/countpairs_rp_pi_mocks.c" "$TMP/gen/countpairs_rp_pi_mocks_impl_float.c" "$TMP/gen/countpairs_rp_pi_mocks_impl_double.c"
  "$REF/mocks/DDsmu_mocks/countpairs_s_mu_mocks.c" "$TMP/gen/countpairs_s_mu_mocks_impl_float.c" "$TMP/gen/countpairs_s_mu_mocks_impl_double.c"
  "$REF/theory/vpf/countspheres.c" "$TMP/gen/countspheres_impl_float.c" "$TMP/gen/countspheres_impl_double.c"
  "$REF/mocks/vpf_mocks/countspheres_mocks.c" "$TMP/gen/countspheres_mocks_impl_float.c" "$TMP/gen/countspheres_mocks_impl_double.c"
  "$REF/utils/cosmology_params.c" "$REF/utils/set_cosmo_dist.c"
  "$TMP/gen/gridlink_impl_float.c" "$TMP/gen/gridlink_impl_double.c"
  "$TMP/gen/gridlink_mocks_impl_float.c" "$TMP/gen/gridlink_mocks_impl_double.c"
  "$TMP/gen/gridlink_utils_float.c" "$TMP/gen/gridlink_utils_double.c"
  "$REF/utils/utils.c" "$REF/utils/progressbar.c" "$REF/utils/cpu_features.c"
)

for variant in v4 v3; do
  mkdir -p "$TMP/obj_$variant"
  objs=()
  for s in "${SRCS[@]}"; do
    o="$TMP/obj_$variant/$(basename "${s%.c}").o"
    # the impl files #include the kernels as "<name>_kernels_<prec>.c" from the gen dir
    gcc $COMMON -march=x86-64-$variant $INCL -c "$s" -o "$o" &
    objs+=("$o")
  done
  wait
  gcc -shared -fopenmp -Wl,-Bsymbolic -o "$OUT/libcorrfunc_ref_$variant.so" "${objs[@]}" -lm
  echo "built $OUT/libcorrfunc_ref_$variant.so"
done
