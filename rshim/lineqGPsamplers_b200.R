# Patched bodies of tmvrnorm.HMC / tmvrnorm.RSM for lineqGPR (dev/lineqGPR/R/lineqGPsamplers.R).
# SOURCE ONLY: R is not installed in the build image, so this file has not been executed here.
# Same function names, same arguments, same d x nsim return value; only the sampling itself
# moves to the GPU through the .Call shim rshim/lqg_rcall.c.
#
# dyn.load("lqg_rcall.so") once (or useDynLib in NAMESPACE).

# c(fork, fork.burn.in) of lqg_opts: with several chains, families of 6 share one root chain's burn-in
# (the root runs burn.in transitions from the clamped MAP; its 3 chains start from the state it reached
# and run 5 more on their own streams before they emit).  control$fork = 1: fully independent chains.
.lqg_fork <- function(control) {
  fork <- if ("fork" %in% names(control)) control$fork else if (control$n.chains > 1) 6L else 1L
  fb <- if ("fork.burn.in" %in% names(control)) control$fork.burn.in else 5L
  as.integer(c(fork, fb))
}

tmvrnorm.HMC <- function(object, nsim, control = list(burn.in = 1e2), ...) {
  if (!("burn.in" %in% names(control))) control$burn.in <- 1e2
  # extension: independent chains.  Default = one chain per resident CTA slot of the device for this
  # dimension (444 on a B200 at d = 2000): every wave full, least burn-in work.  Several chains all
  # start at the clamped MAP, so each of them runs at least the tmvrnorm.HMC default of 100 burn-in
  # transitions before it emits (simulate.* passes burn.in = 1, R/lineqGP.R:264,617, which is meant
  # for ONE long chain); n.chains = 1 is the reference chain and keeps burn.in as given.
  if (!("n.chains" %in% names(control)))
    control$n.chains <- max(1L, min(nsim, .Call("lqg_hmc_box_slots_call", as.integer(length(object$mu)))))
  if (control$n.chains > 1 && !isTRUE(control$burn.in.exact)) control$burn.in <- max(control$burn.in, 100)
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
              as.integer(control$n.chains), as.integer(seed), .lqg_fork(control))
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
  if (identical(control$setup, "host")) {
    # set-up exactly as TruncatedNormal::mvrandn does it (its internal helpers)
    out <- TruncatedNormal:::cholperm(tmvPar$Sigma, l, u)
    Lfull <- out$L; D <- diag(Lfull); perm <- out$perm
    L0 <- Lfull / D - diag(d); l <- out$l / D; u <- out$u / D
    xmu <- TruncatedNormal:::nleq(l, u, L0)
    x <- xmu[1:(d - 1)]; mu <- c(xmu[d:(2 * d - 2)], 0)
    psistar <- TruncatedNormal:::psy(x, L0, l, u, mu)
  } else {
    # the same three steps on the device (permuted Cholesky, Newton tilt, psi*)
    su <- .Call("lqg_expt_setup_call", tmvPar$Sigma, as.double(l), as.double(u))
    L0 <- su[[1]]; l <- su[[2]]; u <- su[[3]]; mu <- su[[4]]; psistar <- su[[5]]; perm <- su[[6]]; Lfull <- su[[7]]
  }
  seed <- sample(1:10000000, 1)
  Z <- .Call("lqg_expt_call", L0, as.double(l), as.double(u), as.double(mu), as.double(psistar),
             as.integer(nsim), as.integer(seed))
  xi <- .Call("lqg_dgemm_call", Lfull[order(perm), , drop = FALSE], Z)
  matrix(tmvPar$mu, nrow = d, ncol = nsim) + xi
}

# ---- simulate.lineqGP / simulate.lineqAGP with the whole tail in one device-resident call ----------
# Same names, arguments and result fields as R/lineqGP.R:601-643 and R/lineqAGP.R:520-596.  The tail
# (tmvrnorm -> qr.solve -> Phi.test %*% xi.sim) goes through .Call("lqg_simulate_call") when the model's
# sampler is "HMC"; any other sampler keeps the reference's three statements (with the GPU tmvrnorm
# methods above).  Extras in the result: ymean / qtls (the bands plot.lineqGP draws, R/lineqGPutils.R:422-423).
.lqg_simulate_tail <- function(tmvPar, Lambda, Phi.test, nsim, control, probs = c(0.025, 0.975), keep.ysim = TRUE) {
  d <- length(tmvPar$mu)
  if (!("map" %in% names(tmvPar))) tmvPar$map <- tmvPar$mu
  if (!is.element(length(tmvPar$lb), c(1, d)) || !is.element(length(tmvPar$ub), c(1, d)))
    stop('The bounds have to be d-dimensionals')
  if (any(tmvPar$lb >= tmvPar$ub))
    stop('The elements from "u" has to be greater than the elements from "l"')
  lb <- rep_len(tmvPar$lb, d); ub <- rep_len(tmvPar$ub, d)
  init_vec <- pmin(pmax(tmvPar$map, lb + 1e-9), ub - 1e-9)          # R/lineqGPsamplers.R:133-135
  if (!("burn.in" %in% names(control))) control$burn.in <- 1e2
  if (!("n.chains" %in% names(control)))
    control$n.chains <- max(1L, min(nsim, .Call("lqg_hmc_box_slots_call", as.integer(d))))
  if (control$n.chains > 1 && !isTRUE(control$burn.in.exact)) control$burn.in <- max(control$burn.in, 100)
  n.per <- ceiling(nsim / control$n.chains)
  seed <- sample(1:10000000, 1)                                     # set.seed(seed) of the caller decides the run
  r <- .Call("lqg_simulate_call", as.double(tmvPar$mu), tmvPar$Sigma, as.double(lb), as.double(ub),
             as.double(init_vec), Lambda, Phi.test, as.double(probs), as.integer(n.per),
             as.integer(control$burn.in), as.integer(control$n.chains), as.integer(seed), isTRUE(keep.ysim),
             .lqg_fork(control))
  keep <- seq_len(nsim)
  list(xi.sim = r[[1]][, keep, drop = FALSE],
       ysim = if (is.null(r[[2]])) NULL else r[[2]][, keep, drop = FALSE],
       ymean = r[[3]], qtls = r[[4]])
}

simulate.lineqGP <- function(object, nsim = 1, seed = NULL, xtest, ...) {
  model <- object
  predtime <- proc.time()
  pred <- predict(model, xtest, return_model = TRUE)
  predtime <- proc.time() - predtime
  model <- pred$model
  pred$model <- NULL
  # eta-space parameters (R/lineqGP.R:611-614) on the GPU
  ep <- .Call("lqg_eta_params_call", pred$Lambda, as.double(pred$mu), as.double(pred$xi.map), pred$Sigma,
              as.double(model$kernParam$nugget))
  control <- as.list(unlist(model$localParam$samplingParam))
  control$mvec <- model$localParam$mvec
  control$constrType <- model$constrType
  tmvPar <- list(mu = ep[[1]], map = ep[[2]], Sigma = ep[[3]], lb = pred$lb, ub = pred$ub)
  class(tmvPar) <- model$localParam$sampler
  set.seed(seed)
  simModel <- list()
  simtime <- proc.time()
  if (identical(model$localParam$sampler, "HMC")) {
    tail <- .lqg_simulate_tail(tmvPar, pred$Lambda, pred$Phi.test, nsim, control)
    simModel$xi.sim <- tail$xi.sim
    simModel$ysim <- tail$ysim
    simModel$ymean <- tail$ymean; simModel$qtls <- tail$qtls
  } else {
    eta <- tmvrnorm(tmvPar, nsim, control)
    simModel$xi.sim <- .Call("lqg_dgemm_call", .Call("lqg_lambda_pinv_call", pred$Lambda), eta)   # qr.solve(pred$Lambda, eta)
    simModel$ysim <- .Call("lqg_dgemm_call", pred$Phi.test, simModel$xi.sim)
  }
  simtime <- proc.time() - simtime
  simModel$x <- model$x
  simModel$y <- model$y
  simModel$xtest <- xtest
  simModel$Phi.test <- pred$Phi.test
  simModel$predtime <- predtime
  simModel$simtime <- simtime
  simModel$xi.map <- pred$xi.map
  class(simModel) <- class(model)
  return(simModel)
}

simulate.lineqAGP <- function(object, nsim = 1, seed = NULL, xtest, ...) {
  model <- object
  predtime <- proc.time()
  pred <- predict(model, xtest, return_model = TRUE)
  predtime <- proc.time() - predtime
  model <- pred$model
  pred$model <- NULL
  LambdaAll <- as.matrix(model$lineqSys$LambdaAll)
  ep <- .Call("lqg_eta_params_call", LambdaAll, as.double(pred$xiAll.map), as.double(pred$xiAll.map),
              as.matrix(pred$SigmaAll), as.double(model$nugget))          # R/lineqAGP.R:558-560
  control <- as.list(unlist(model$localParam$samplingParam))
  control$mvec <- model$localParam$mtotal
  control$constrType <- model$constrType
  tmvPar <- list(mu = ep[[2]], Sigma = ep[[3]], lb = model$lineqSys$lb, ub = model$lineqSys$ub)   # mu := eta.map (:567)
  class(tmvPar) <- model$localParam$sampler
  set.seed(seed)
  simtime <- proc.time()
  if (identical(model$localParam$sampler, "HMC")) {
    xiAll.sim <- .lqg_simulate_tail(tmvPar, LambdaAll, NULL, nsim, control)$xi.sim
  } else {
    etaAll <- tmvrnorm(tmvPar, nsim, control)
    xiAll.sim <- .Call("lqg_dgemm_call", .Call("lqg_lambda_pinv_call", LambdaAll), etaAll)    # qr.solve(LambdaAll, etaAll)
  }
  simtime <- proc.time() - simtime
  simModel <- list()
  simModel$x <- model$x
  simModel$y <- model$y
  simModel$xtest <- xtest
  simModel$Phi.test <- pred$Phi.test
  simModel$predtime <- predtime
  simModel$simtime <- simtime
  simModel$muAll <- pred$muAll
  simModel$xiAll.map <- pred$xiAll.map
  simModel$xiAll.sim <- xiAll.sim
  simModel$PhiAll.test <- pred$PhiAll.test
  class(simModel) <- class(model)
  return(simModel)
}

# The rest of the set-up stage of predict.lineqGP on the GPU (rshim/lqg_rcall.c, second half):
#   R/lineqGP.R:518   invPhiGammaPhit <- chol2inv(chol(PhiGammaPhit))     -> .Call("lqg_solve_spd_call", PhiGammaPhit, diag(n))
#   R/lineqGP.R:527   if (min(eigen(pred$Sigma)$values) <= 0)             -> if (!.Call("lqg_is_pd_call", pred$Sigma))
#   R/lineqGP.R:529   invSigma <- chol2inv(chol(pred$Sigma))              -> .Call("lqg_solve_spd_call", pred$Sigma, diag(m))
#   R/lineqGP.R:530   solve.QP(invSigma, t(mu) %*% invSigma, t(M), g)$solution
#                                                                         -> .Call("lqg_solve_qp_call", invSigma, mu %*% invSigma, M, g)
# paths and bands of an existing xi.sim, without keeping the n_t x N matrix (config 5):
#   r <- .Call("lqg_paths_call", PhiAll.test, xiAll.sim, probs, FALSE); ymean <- r[[2]]; qtls <- r[[3]]
# plot.lineqGP bands (R/lineqGPutils.R:422-423):
#   b <- .Call("lqg_bands_call", ysim, probs); qtls <- b[[2]]; ymean <- b[[1]]
