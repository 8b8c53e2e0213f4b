/* TEST INFRASTRUCTURE ONLY -- stand-in for <gsl/gsl_integration.h> (GSL is not installed in this image).
 * utils/set_cosmo_dist.c includes this header but calls nothing from it: its redshift -> distance table is a plain
 * Simpson loop.  With this empty header oracle/build_ref.sh compiles that file UNMODIFIED. */
#pragma once
