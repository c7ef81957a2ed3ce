# Host side of GPC (reference: R/GPCclass.R:55-210 of the CPU package).  The Newton loop of the Laplace approximation
# runs in libgprc; its stopping rule and the reference's divergence guard are evaluated there with the reference's
# exact comparisons and reported back as `status`.  The per-point quadrature stays in R (integrate()).

#' @export
GPC <- R6::R6Class("GPC",
  private = list(.X = NULL, .k = NULL, .y = NULL, .logq = NULL, .ptr = NULL,
    .sigmoid = function(x) 1 / (1 + exp(-x)),
    readonly = function(name, value, missing_value) {
      if (!missing_value) stop(sprintf("`$%s` is read only", name), call. = FALSE)
    }),
  public = list(
    initialize = function(X, y, k, epsilon = 1e-5) {
      stopifnot(is.numeric(X), is.vector(y), is.numeric(y))
      stopifnot(is.numeric(epsilon), epsilon > 0, is.function(k))
      if (!is.matrix(X)) dim(X) <- c(1, length(X))
      stopifnot(length(y) == ncol(X))
      storage.mode(X) <- "double"
      spec <- .gprc_spec(k)
      K <- if (is.null(spec)) covariance_matrix(X, X, k) else NULL
      res <- .Call(C_gprc_gpc_fit, spec, X, as.double(y), as.double(epsilon), TRUE, K)
      status <- res[[6]]
      if (status == 1L) stop("Apparently does not converge.")
      if (status != 0L) stop("the leading minor of I + W^(1/2) K W^(1/2) is not positive definite")
      message(sprintf("Convergence after %s iterations", res[[2]]))
      trace <- res[[3]]
      private$.ptr <- res[[1]]
      private$.logq <- trace[length(trace)] - res[[4]]      # objective - sum(diag(L)), as the CPU package computes it
      private$.X <- X; private$.y <- y; private$.k <- k
    },
    predict_class = function(X_star) {
      if (!is.matrix(X_star)) dim(X_star) <- c(1, length(X_star))
      if (nrow(X_star) != nrow(private$.X)) stop("non-conformable arrays")   # as covariance_matrix would (R/GPRclass.R:356)
      storage.mode(X_star) <- "double"
      spec <- .gprc_spec(private$.k)
      # built-in kernels: latent prediction and the integrate() loop in one device call (dqagi port, same tolerances)
      if (!is.null(spec) && !isFALSE(getOption("gprc.device_quadrature")))
        return(.Call(C_gprc_gpc_predict_class, private$.ptr, X_star))
      lat <- if (!is.null(spec)) .Call(C_gprc_gpc_predict_latent, private$.ptr, X_star, NULL, NULL)
             else .Call(C_gprc_gpc_predict_latent, private$.ptr, NULL, covariance_matrix(private$.X, X_star, private$.k),
                        as.double(private$.k(X_star, X_star)))
      # column 1 = fs_bar, column 2 = Vfs; the latent variance is passed as `sd`, like the CPU package does
      vapply(seq_len(nrow(lat)), function(i)
        integrate(function(z) private$.sigmoid(z) * dnorm(z, mean = lat[i, 1], sd = lat[i, 2]), -Inf, Inf)$value,
        numeric(1))
    }
  ),
  active = list(
    X = function(value) { private$readonly("X", value, missing(value)); private$.X },
    k = function(value) { private$readonly("k", value, missing(value)); private$.k },
    y = function(value) { private$readonly("y", value, missing(value)); private$.y },
    logq = function(value) { private$readonly("logq", value, missing(value)); private$.logq },
    f_hat = function(value) { private$readonly("f_hat", value, missing(value)); .Call(C_gprc_gpc_get, private$.ptr, 3L) },
    L = function(value) { private$readonly("L", value, missing(value)); .Call(C_gprc_gpc_get, private$.ptr, 0L) }
  )
)
