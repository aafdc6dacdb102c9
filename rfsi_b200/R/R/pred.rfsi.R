# rfsi / pred.rfsi -- the wrappers of README.md:93-143 over the engine.
# rfsi(): near.obs on the GPU for the training table, ranger() itself for training (training is
# out of scope for the engine), then the forest is imported.  pred.rfsi(): one fused
# near.obs -> forest call for all prediction locations and days.
rfsi <- function(formula, data, data.staid.x.y.z = NULL, n.obs = 5, cpus = 1, progress = TRUE, zero.tol = 0, ...) {
  if (zero.tol != 0) stop("rfsi (rfsi_b200): only zero.tol = 0 is supported")
  d <- as.data.frame(data); c4 <- data.staid.x.y.z
  days <- if (is.null(c4) || is.na(c4[4])) rep(1L, nrow(d)) else d[[c4[4]]]
  zname <- all.vars(formula)[1]
  feats <- do.call(rbind, lapply(split(seq_len(nrow(d)), days), function(i) {
    no <- near.obs(locations = d[i, ], locations.x.y = c4[2:3], observations = d[i, ],
                   observations.x.y = c4[2:3], zcol = zname, n.obs = n.obs)
    cbind(d[i, ], no)
  }))
  feats <- feats[stats::complete.cases(feats), ]
  fm <- stats::update(formula, paste(". ~ . +", paste(c(paste0("dist", 1:n.obs), paste0("obs", 1:n.obs)), collapse = "+")))
  model <- ranger::ranger(fm, data = feats, ...)
  model$rfsi_b200 <- rfsi_import_ranger(model); model$n.obs <- n.obs
  model
}
pred.rfsi <- function(model, data, obs.col = 1, data.staid.x.y.z = NULL, newdata, newdata.staid.x.y.z = NULL,
                      output.format = "data.frame", zero.tol = 0, cpus = 1, progress = TRUE, soil3d = FALSE,
                      type = "response", quantiles = c(0.1, 0.5, 0.9), device = 0L, ...) {
  if (zero.tol != 0 || soil3d) stop("pred.rfsi (rfsi_b200): zero.tol = 0 and soil3d = FALSE only")
  fo <- if (is.null(model$rfsi_b200)) rfsi_import_ranger(model) else model$rfsi_b200
  n.obs <- sum(grepl("^dist[0-9]+$", fo$independent.variable.names))
  d <- as.data.frame(data); nd <- as.data.frame(newdata); c4 <- data.staid.x.y.z; n4 <- newdata.staid.x.y.z
  sta <- unique(d[, c4[1:3]]); dayv <- if (is.na(c4[4])) 1L else sort(unique(d[[c4[4]]]))
  z <- matrix(NA_real_, nrow(sta), length(dayv))            # stations x days, NA = no observation
  z[cbind(match(d[[c4[1]]], sta[[1]]), if (is.na(c4[4])) 1L else match(d[[c4[4]]], dayv))] <- d[[obs.col]]
  covn <- setdiff(fo$independent.variable.names, c(paste0("dist", 1:n.obs), paste0("obs", 1:n.obs)))
  code <- vapply(fo$independent.variable.names, function(v) {
    if (grepl("^dist[0-9]+$", v)) 65536L + as.integer(sub("dist", "", v)) - 1L
    else if (grepl("^obs[0-9]+$", v)) 131072L + as.integer(sub("obs", "", v)) - 1L
    else match(v, covn) - 1L }, 1L)
  res <- .Call(C_rfsi_pred_fused, fo$handle, as.matrix(nd[, n4[2:3]]) + 0, as.matrix(sta[, 2:3]) + 0, z,
               as.integer(n.obs), lapply(covn, function(v) as.double(nd[[v]])), as.integer(code),
               if (type == "quantiles") as.double(quantiles) else NULL, as.integer(device))
  out <- data.frame(nd[, n4[1:3]], pred = as.vector(res[[1]]))
  if (type == "quantiles") {
    q <- matrix(res[[2]], ncol = length(quantiles)); colnames(q) <- paste0("quantile..", quantiles)
    out <- cbind(out, q)
  }
  out
}
