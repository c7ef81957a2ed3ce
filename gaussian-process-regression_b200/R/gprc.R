#' gprc: Gaussian-process regression and classification on the GPU
#'
#' Same exported API as the CPU package (GPR, GPC, fit, cov_func, iid_noise, simulate_*); the numerical hot path
#' (kernel-matrix build, Cholesky, triangular solves, Laplace Newton loop) runs in libgprc through `.Call`.
#' Device selection: options(gprc.device = 0L).  There is no CPU fallback.
#'
#' @useDynLib gprc, .registration = TRUE
#' @importFrom R6 R6Class
"_PACKAGE"

# ---- kernel descriptors ------------------------------------------------------------------------------------------
# A closure made by cov_func()/fit() carries attr(k, "gprc_kernel") = list(id=, <named parameters>); GPR/GPC hand such
# kernels to the device.  Any other function(x, y) is evaluated in R exactly as before (covariance_matrix via outer)
# and only the factorisation / solves run on the device ("precomputed" entry points).
.gprc_ids <- c(constant = 0L, linear = 1L, polynomial = 2L, sqrexp = 3L, gammaexp = 4L, rationalquadratic = 5L)

.gprc_tag <- function(closure, func, args) {
  for (nm in names(.gprc_ids)) {
    if (identical(func, get(nm, envir = asNamespace("gprc")))) {
      formal <- setdiff(names(formals(get(paste0(nm, ".matrix"), envir = asNamespace("gprc")))), c("x", "y"))
      if (is.null(names(args))) names(args) <- rep("", length(args))
      unnamed <- which(names(args) == "")
      names(args)[unnamed] <- setdiff(formal, names(args))[seq_along(unnamed)]
      if (setequal(names(args), formal))
        attr(closure, "gprc_kernel") <- c(list(id = .gprc_ids[[nm]]), lapply(args[formal], as.double))
    }
  }
  closure
}

.gprc_spec <- function(k) attr(k, "gprc_kernel", exact = TRUE)

# ---- library options (include/gprc.h, GPRC_OPT_*) ---------------------------------------------------------------------
# gprc_options(predict_path = 0:4, ozaki_digits = 6:8, int8_auto = 0:1, int8_tile = c(2, 1, 64, 128), gram_dmma = 0:2,
#              trsv = 0:1, chol_tiles = 0:1000)
# predict_path: 0 automatic; 1 inverse + triangular GEMM; 2 / 3 FP64 blocked substitution; 4 the substitution with its
# products on the INT8 tensor cores (the automatic choice for >= 18 944 test points and n >= 4096).
#' @export
gprc_options <- function(...) {
  ids <- c(gram_dmma = 1L, predict_path = 2L, ozaki_digits = 3L, int8_auto = 4L, int8_tile = 5L, trsv = 7L,
           chol_tiles = 8L)
  args <- list(...)
  stopifnot(all(names(args) %in% names(ids)))
  for (nm in names(args)) .Call(C_gprc_set_option, ids[[nm]], as.integer(args[[nm]]))
  invisible(NULL)
}
