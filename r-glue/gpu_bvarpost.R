# gpu_bvarpost.R -- the FUN argument of draw_posterior() (R/draw_posterior.bvarmodel.R:39,71-86):
#
#   object <- draw_posterior(model, FUN = gpu_bvarpost)           # per model: one chain on the GPU, a `bvar` object
#   chains <- gpu_bvarpost(model, n_chains = 4096, n_gpus = -1)   # 4096 chains of one model over all GPUs: a bvarlist
#
# bvarpost()'s own post-processing (R/bvarpost.R:64-119: A0 / sigma-lambda reshaping, bvar()) is reused unchanged: only
# the call into C++ differs.  Do not combine with mc.cores: forked children must not inherit a CUDA context; the library
# shards chains over the GPUs itself (n_gpus).
gpu_bvarpost <- function(object, n_chains = 1L, n_gpus = 1L, device = -1L, seed = -1) {
  alg <- if (object[["model"]][["tvp"]]) .bvartvpalg_gpu else .bvaralg_gpu
  res <- alg(object, as.integer(n_chains), as.integer(n_gpus), as.integer(device), as.numeric(seed))
  finish <- function(obj) {
    # R/bvarpost.R:64-119 verbatim from here on (A0, sigma$lambda, bvar())
    if (!is.null(obj[["posteriors"]][["sigma"]][["lambda"]])) {
      k <- NCOL(obj[["data"]][["Y"]])
      sigma_lambda <- matrix(diag(NA_real_, k), k * k, obj[["model"]][["iterations"]])
      sigma_lambda[which(lower.tri(diag(1, k))), ] <- obj[["posteriors"]][["sigma"]][["lambda"]]
      sigma_lambda[which(upper.tri(diag(1, k))), ] <- obj[["posteriors"]][["sigma"]][["lambda"]]
      obj[["posteriors"]][["sigma"]][["lambda"]] <- sigma_lambda
    }
    A0 <- NULL
    if (obj[["model"]][["structural"]]) {
      k <- NCOL(obj[["data"]][["Y"]]); tt <- NROW(obj[["data"]][["Y"]])
      pos <- which(lower.tri(diag(1, k))); draws <- obj[["model"]][["iterations"]]
      a0 <- obj[["posteriors"]][["a0"]]
      if (obj[["model"]][["tvp"]]) {
        A0[["coeffs"]] <- matrix(diag(1, k), k * k * tt, draws)
        A0[["coeffs"]][rep(0:(tt - 1) * k * k, each = length(pos)) + rep(pos, tt), ] <- a0[["coeffs"]]
      } else {
        A0[["coeffs"]] <- matrix(diag(1, k), k * k, draws)
        A0[["coeffs"]][rep(0:(draws - 1) * k * k, each = length(pos)) + pos] <- a0[["coeffs"]]
      }
      if ("sigma" %in% names(a0)) { A0[["sigma"]] <- matrix(0, k * k, draws); A0[["sigma"]][pos, ] <- a0[["sigma"]] }
      if ("lambda" %in% names(a0)) {
        A0[["lambda"]] <- matrix(diag(1, k), k * k, draws)
        A0[["lambda"]][pos, ] <- a0[["lambda"]]; A0[["lambda"]][-pos, ] <- NA_real_
      }
    }
    bvartools::bvar(data = NULL, exogen = NULL, y = obj[["data"]][["Y"]], x = obj[["data"]][["Z"]], A0 = A0,
                    A = obj[["posteriors"]][["a"]], B = obj[["posteriors"]][["b"]], C = obj[["posteriors"]][["c"]],
                    Sigma = obj[["posteriors"]][["sigma"]])
  }
  if (n_chains == 1L) return(finish(res))
  out <- lapply(res, finish)
  class(out) <- append("bvarlist", class(out))
  out
}
