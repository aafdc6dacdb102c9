# rfsi / pred.rfsi -- the wrappers of README.md:93-143 over the engine, with the argument lists
# and input classes of that example: `data` an sf / SpatVector / data.frame, `newdata` an
# sf / SpatVector / SpatRaster / data.frame, `output.format` "data.frame" | "sf" | "SpatVector" |
# "SpatRaster".  (R is not installed where this repository is built: these files are exercised by
# rfsi_b200/R/tests/parity_against_meteo_ranger.R on a maintainer's machine; the .Call entries they
# use are run in the test-suite through a mock of R's C API.)
#
# rfsi(): near.obs on the GPU for the training table, ranger() itself for training (training is
# out of scope for the engine), then the forest is imported.  pred.rfsi(): one fused
# near.obs -> forest call per time step: every newdata row meets the observations of ITS day.

# Any of the supported classes -> list(df = plain data.frame incl. id / x / y (/ time) columns,
# cols = c(staid, x, y, time-or-NA) as column NAMES, template = SpatRaster or NULL, crs)
.rfsi_table <- function(obj, staid.x.y.z, what) {
  if (inherits(obj, "SpatRaster")) {
    df <- terra::as.data.frame(obj, xy = TRUE, cells = TRUE, na.rm = FALSE)
    names(df)[1:3] <- c(".cell", ".x", ".y")
    return(list(df = df, cols = c(".cell", ".x", ".y", NA), template = obj, crs = terra::crs(obj)))
  }
  if (inherits(obj, "sf")) {
    xy <- sf::st_coordinates(obj)
    df <- sf::st_drop_geometry(obj)
    df$.id <- seq_len(nrow(df)); df$.x <- xy[, 1]; df$.y <- xy[, 2]
    tcol <- if (inherits(obj, "sftime")) attr(obj, "time_column") else NA
    return(list(df = as.data.frame(df), cols = c(".id", ".x", ".y", tcol), template = NULL, crs = sf::st_crs(obj)))
  }
  if (inherits(obj, "SpatVector")) {
    xy <- terra::crds(obj)
    df <- terra::as.data.frame(obj)
    df$.id <- seq_len(nrow(df)); df$.x <- xy[, 1]; df$.y <- xy[, 2]
    return(list(df = df, cols = c(".id", ".x", ".y", NA), template = NULL, crs = terra::crs(obj)))
  }
  if (inherits(obj, "Spatial")) {
    xy <- sp::coordinates(obj)
    df <- as.data.frame(obj@data)
    df$.id <- seq_len(nrow(df)); df$.x <- xy[, 1]; df$.y <- xy[, 2]
    return(list(df = df, cols = c(".id", ".x", ".y", NA), template = NULL, crs = NA))
  }
  df <- as.data.frame(obj)
  if (is.null(staid.x.y.z))
    stop(what, ": a data.frame needs ", what, ".staid.x.y.z = c(<id>, <x>, <y>, <time or NA>)")
  cols <- staid.x.y.z
  if (is.numeric(cols)) cols <- ifelse(is.na(cols), NA_character_, names(df)[cols])
  if (length(cols) < 4) cols <- c(cols, NA)
  list(df = df, cols = cols, template = NULL, crs = NA)
}

rfsi <- function(formula, data, data.staid.x.y.z = NULL, n.obs = 5, avg = FALSE, increment, range, quadrant = FALSE,
                 use.idw = FALSE, idw.p = 2, s.crs = NA, p.crs = NA, cpus = 1, progress = TRUE, soil3d = FALSE,
                 depth.range = 0.1, no.obs = "increase", zero.tol = 0, ...) {
  if (avg || quadrant || use.idw || soil3d) stop("rfsi (rfsi_b200): avg / quadrant / use.idw / soil3d are not supported")
  if (zero.tol != 0) stop("rfsi (rfsi_b200): only zero.tol = 0 is supported")
  tb <- .rfsi_table(data, data.staid.x.y.z, "data")
  d <- tb$df; c4 <- tb$cols
  days <- if (is.na(c4[4])) rep(1L, nrow(d)) else d[[c4[4]]]
  zname <- all.vars(formula)[1]
  feats <- do.call(rbind, lapply(split(seq_len(nrow(d)), days), function(i) {
    no <- near.obs(locations = d[i, ], locations.x.y = c4[2:3], observations = d[i, ],
                   observations.x.y = c4[2:3], zcol = zname, n.obs = n.obs)
    cbind(d[i, ], no)
  }))
  feats <- feats[stats::complete.cases(feats[, c(all.vars(formula), names(feats)[grepl("^(dist|obs)[0-9]+$", names(feats))])]), ]
  fm <- stats::update(formula, paste(". ~ . +", paste(c(paste0("dist", 1:n.obs), paste0("obs", 1:n.obs)), collapse = "+")))
  model <- ranger::ranger(fm, data = feats, ...)
  model$n.obs <- n.obs
  model
}

pred.rfsi <- function(model, data, obs.col = 1, data.staid.x.y.z = NULL, newdata, newdata.staid.x.y.z = NULL,
                      output.format = "data.frame", zero.tol = 0, s.crs = NA, newdata.s.crs = NA, p.crs = NA,
                      cpus = 1, progress = TRUE, soil3d = FALSE, depth.range = 0.1, no.obs = "increase",
                      type = "response", quantiles = c(0.1, 0.5, 0.9), device = 0L, ...) {
  if (zero.tol != 0 || soil3d) stop("pred.rfsi (rfsi_b200): zero.tol = 0 and soil3d = FALSE only")
  if (!is.na(s.crs) || !is.na(p.crs) || !is.na(newdata.s.crs))
    stop("pred.rfsi (rfsi_b200): re-projection (s.crs / p.crs) is not done here; pass projected coordinates")
  fo <- rfsi_forest_of(model)
  vars <- fo$independent.variable.names
  n.obs <- sum(grepl("^dist[0-9]+$", vars))
  ob <- .rfsi_table(data, data.staid.x.y.z, "data"); d <- ob$df; c4 <- ob$cols
  nw <- .rfsi_table(newdata, newdata.staid.x.y.z, "newdata"); nd <- nw$df; n4 <- nw$cols
  if (is.numeric(obs.col)) obs.col <- names(as.data.frame(data))[obs.col]

  # stations x days table of observations, NA = no observation that day (the scripts' !is.na filter)
  sta <- d[!duplicated(d[[c4[1]]]), c4[1:3]]
  dayv <- if (is.na(c4[4])) 1L else sort(unique(d[[c4[4]]]))
  z <- matrix(NA_real_, nrow(sta), length(dayv))
  z[cbind(match(d[[c4[1]]], sta[[1]]), if (is.na(c4[4])) 1L else match(d[[c4[4]]], dayv))] <- as.double(d[[obs.col]])

  covn <- setdiff(vars, c(paste0("dist", seq_len(n.obs)), paste0("obs", seq_len(n.obs))))
  miss <- setdiff(covn, names(nd))
  if (length(miss)) stop("pred.rfsi: covariates missing from newdata: ", paste(miss, collapse = ", "))
  code <- vapply(vars, function(v) {
    if (grepl("^dist[0-9]+$", v)) 65536L + as.integer(sub("dist", "", v)) - 1L
    else if (grepl("^obs[0-9]+$", v)) 131072L + as.integer(sub("obs", "", v)) - 1L
    else match(v, covn) - 1L }, 1L)
  probs <- if (type == "quantiles") as.double(quantiles) else NULL

  # rows whose covariates are complete (a raster's NA cells stay NA in the output)
  ok <- if (length(covn)) stats::complete.cases(nd[, covn, drop = FALSE]) else rep(TRUE, nrow(nd))
  nday <- if (is.na(n4[4])) rep(dayv[1], nrow(nd)) else nd[[n4[4]]]
  if (is.na(n4[4]) && length(dayv) > 1)
    stop("pred.rfsi: data has ", length(dayv), " time steps but newdata has no time column (newdata.staid.x.y.z[4])")
  pred <- rep(NA_real_, nrow(nd))
  q <- if (is.null(probs)) NULL else matrix(NA_real_, nrow(nd), length(probs))
  for (dd in unique(nday[ok])) {                       # every newdata row meets the observations of its own day
    k <- match(dd, dayv)
    if (is.na(k)) stop("pred.rfsi: newdata has a time step without observations: ", format(dd))
    i <- which(ok & nday == dd)
    loc <- cbind(as.double(nd[[n4[2]]][i]), as.double(nd[[n4[3]]][i]))
    res <- .Call(C_rfsi_pred_fused, fo$handle, loc, cbind(as.double(sta[[2]]), as.double(sta[[3]])),
                 z[, k, drop = FALSE], as.integer(n.obs), lapply(covn, function(v) as.double(nd[[v]][i])),
                 as.integer(code), probs, as.integer(device))
    pred[i] <- as.vector(res[[1]])
    if (!is.null(probs)) q[i, ] <- matrix(res[[2]], ncol = length(probs))
  }

  out <- data.frame(nd[, n4[!is.na(n4)], drop = FALSE], pred = pred)
  if (!is.null(probs)) { colnames(q) <- paste0("quantile..", quantiles); out <- cbind(out, q) }
  valn <- setdiff(names(out), n4[!is.na(n4)])
  if (output.format == "SpatRaster") {
    if (is.null(nw$template)) stop("pred.rfsi: output.format = 'SpatRaster' needs a SpatRaster newdata")
    r <- terra::rast(nw$template, nlyrs = length(valn))
    terra::values(r) <- as.matrix(out[order(nd[[n4[1]]]), valn, drop = FALSE])
    names(r) <- valn
    return(r)
  }
  names(out)[match(n4[1:3], names(out))] <- c("staid", "x", "y")
  if (output.format == "sf") return(sf::st_as_sf(out, coords = c("x", "y"), crs = nw$crs, agr = "constant"))
  if (output.format == "SpatVector") return(terra::vect(out, geom = c("x", "y"), crs = if (is.na(nw$crs)[1]) "" else nw$crs))
  out
}
