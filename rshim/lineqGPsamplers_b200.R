# Patched bodies of tmvrnorm.HMC / tmvrnorm.RSM for lineqGPR (dev/lineqGPR/R/lineqGPsamplers.R).
# SOURCE ONLY: R is not installed in the build image, so this file has not been executed here.
# Same function names, same arguments, same d x nsim return value; only the sampling itself
# moves to the GPU through the .Call shim rshim/lqg_rcall.c.
#
# dyn.load("lqg_rcall.so") once (or useDynLib in NAMESPACE).

tmvrnorm.HMC <- function(object, nsim, control = list(burn.in = 1e2), ...) {
  if (!("burn.in" %in% names(control))) control$burn.in <- 1e2
  if (!("n.chains" %in% names(control))) control$n.chains <- min(nsim, 1024L)   # extension
  tmvPar <- object
  if (!("map" %in% names(tmvPar))) tmvPar$map <- tmvPar$mu
  d <- length(tmvPar$mu)
  # the checks of bounds2lineqSys(d, lb, ub, A = diag(d), "oneside")  (mvec := length(mu), see Q3)
  if (!is.element(length(tmvPar$lb), c(1, d)) || !is.element(length(tmvPar$ub), c(1, d)))
    stop('The bounds have to be d-dimensionals')
  if (any(tmvPar$lb >= tmvPar$ub))
    stop('The elements from "u" has to be greater than the elements from "l"')
  lb <- rep_len(tmvPar$lb, d); ub <- rep_len(tmvPar$ub, d)
  epsilon <- 1e-9
  init_vec <- pmin(pmax(tmvPar$map, lb + epsilon), ub - epsilon)
  # whitening: the Ri of tmg::rtmg (tmg/R/tmg.R:15-27: H <- solve(Sigma); M <- (H+t(H))/2; eigen test;
  # Ri <- solve(chol(M))) is the unique upper-triangular factor with Ri %*% t(Ri) = Sigma; the device
  # computes it directly (reversed blocked Cholesky) and stops with tmg's message when M is not
  # positive definite.  control$whiten = "host" keeps the literal R sequence.
  if (identical(control$whiten, "host")) {
    H <- solve(tmvPar$Sigma)
    M <- (H + t(H)) / 2
    if (any(eigen(M, symmetric = TRUE, only.values = TRUE)$values <= 0))
      stop("Error: M must be positive definite.")
    Ri <- solve(chol(M))
  } else {
    Ri <- .Call("lqg_whiten_box_call", tmvPar$Sigma)
  }
  n.per <- ceiling(nsim / control$n.chains)
  seed <- sample(1:10000000, 1)             # tmg/R/tmg.R:102: set.seed() still determines the run
  xi <- .Call("lqg_hmc_box_call", as.double(tmvPar$mu), Ri, tmvPar$Sigma, as.double(lb), as.double(ub),
              as.double(init_vec), as.integer(n.per), as.integer(control$burn.in),
              as.integer(control$n.chains), as.integer(seed))
  return(xi[, seq_len(nsim), drop = FALSE])
}

tmvrnorm.RSM <- function(object, nsim, control = NULL, ...) {
  tmvPar <- object
  if (!("map" %in% names(tmvPar))) tmvPar$map <- tmvPar$mu
  invSigma <- chol2inv(chol(tmvPar$Sigma))
  w <- as.vector(invSigma %*% tmvPar$map)
  eS <- eigen(tmvPar$Sigma, symmetric = TRUE)            # MASS::mvrnorm's factor, hoisted out of the loop
  fac <- eS$vectors %*% diag(sqrt(pmax(eS$values, 0)), length(w))
  d <- length(w)
  seed <- sample(1:10000000, 1)
  .Call("lqg_rsm_call", as.double(tmvPar$map), fac, w, as.double(rep_len(tmvPar$lb, d)),
        as.double(rep_len(tmvPar$ub, d)), as.integer(nsim), as.integer(seed))
}

tmvrnorm.ExpT <- function(object, nsim, control = NULL, ...) {
  tmvPar <- object
  d <- length(tmvPar$mu)
  l <- tmvPar$lb - tmvPar$mu; u <- tmvPar$ub - tmvPar$mu
  # set-up exactly as TruncatedNormal::mvrandn does it (its internal helpers)
  out <- TruncatedNormal:::cholperm(tmvPar$Sigma, l, u)
  Lfull <- out$L; D <- diag(Lfull); perm <- out$perm
  L0 <- Lfull / D - diag(d); l <- out$l / D; u <- out$u / D
  xmu <- TruncatedNormal:::nleq(l, u, L0)
  x <- xmu[1:(d - 1)]; mu <- c(xmu[d:(2 * d - 2)], 0)
  psistar <- TruncatedNormal:::psy(x, L0, l, u, mu)
  seed <- sample(1:10000000, 1)
  Z <- .Call("lqg_expt_call", L0, as.double(l), as.double(u), as.double(mu), as.double(psistar),
             as.integer(nsim), as.integer(seed))
  xi <- .Call("lqg_dgemm_call", Lfull[order(perm), , drop = FALSE], Z)
  matrix(tmvPar$mu, nrow = d, ncol = nsim) + xi
}

# simulate.lineqGP tail (R/lineqGP.R:639-640) on the GPU:
#   simModel$xi.sim <- .Call("lqg_dgemm_call", qr.solve(pred$Lambda, diag(nrow(pred$Lambda))), eta)
#   simModel$ysim   <- .Call("lqg_dgemm_call", pred$Phi.test, simModel$xi.sim)
# set-up stage of predict.lineqGP / simulate.lineqGP on the GPU (rshim/lqg_rcall.c, second half):
#   R/lineqGP.R:518   invPhiGammaPhit <- chol2inv(chol(PhiGammaPhit))     -> .Call("lqg_solve_spd_call", PhiGammaPhit, diag(n))
#   R/lineqGP.R:527   if (min(eigen(pred$Sigma)$values) <= 0)             -> if (!.Call("lqg_is_pd_call", pred$Sigma))
#   R/lineqGP.R:529   invSigma <- chol2inv(chol(pred$Sigma))              -> .Call("lqg_solve_spd_call", pred$Sigma, diag(m))
#   R/lineqGP.R:530   solve.QP(invSigma, t(mu) %*% invSigma, t(M), g)$solution
#                                                                         -> .Call("lqg_solve_qp_call", invSigma, mu %*% invSigma, M, g)
#   R/lineqGP.R:611-614  eta.mu / eta.map / Sigma.eta                     -> .Call("lqg_eta_params_call", Lambda, mu, xi.map, Sigma, nugget)
#   R/lineqGP.R:639   qr.solve(pred$Lambda, eta)                          -> .Call("lqg_dgemm_call", .Call("lqg_lambda_pinv_call", pred$Lambda), eta)
# paths and bands in one call, without keeping the n_t x N matrix (config 5):
#   r <- .Call("lqg_paths_call", pred$Phi.test, simModel$xi.sim, probs, FALSE); ymean <- r[[2]]; qtls <- r[[3]]
# plot.lineqGP bands (R/lineqGPutils.R:422-423):
#   b <- .Call("lqg_bands_call", ysim, probs); qtls <- b[[2]]; ymean <- b[[1]]
