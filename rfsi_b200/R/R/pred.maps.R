# pred.maps -- the map-making loop of prcp_catalonia/5_prcp_prediction.R:244-309 in one engine call.
#
#   gg   : prediction locations in the covariate rasters' CRS (what extract() is given), npix x 2
#   gg2  : the same locations in the CRS near.obs works in (spTransform(gg, utm31)), npix x 2
#   obs.xy, obs.z : station coordinates (nobs x 2, CRS of gg2) and the nobs x ndays table of
#                   observed values (NA = no observation that day; days with < 2 stations: drop before)
#   covariates : named list in the model's covariate names; each element a numeric vector (npix),
#                a matrix (npix x ndays), or a Raster* (package raster) with one layer (static) or
#                one layer per day -- sampled on the GPU like extract(r, gg)
#   divisor    : named numeric, e.g. c(tmin = 10, tmax = 10) for `extract(tmin, gg) / 10`
#   cov.fix    : list(c("tmin", "tmax", 1)) for `df$tmin <- ifelse(df$tmin > df$tmax, df$tmax-1, df$tmin)`
#   device     : one GPU index, or several (0:7): the pixels are cut into one band per GPU inside
#                the library (rfsi_predict_fused_multi) -- one R process drives the whole box
# Returns list(pred, iqr): integer npix x ndays matrices, round(x*100), NA where the scripts
# write NAflag = -32767; hand them to writeRaster(..., datatype='INT2S') as before.
pred.maps <- function(model, gg, gg2 = gg, obs.xy, obs.z, n.obs, covariates, divisor = NULL, cov.fix = NULL,
                      quantiles = c(0.25, 0.5, 0.75), device = 0L) {
  fo <- rfsi_forest_of(model)
  covn <- names(covariates)
  code <- vapply(fo$independent.variable.names, function(v) {
    if (grepl("^dist[0-9]+$", v)) 65536L + as.integer(sub("dist", "", v)) - 1L
    else if (grepl("^obs[0-9]+$", v)) 131072L + as.integer(sub("obs", "", v)) - 1L
    else if (v %in% covn) match(v, covn) - 1L
    else stop("pred.maps: model column ", v, " is neither dist/obs nor a supplied covariate") }, 1L)
  cov <- lapply(covn, function(n) {
    r <- covariates[[n]]
    if (inherits(r, "Raster")) {
      v <- raster::getValues(r)                       # cells row by row from the top-left; NAvalue -> NA
      if (!is.integer(v)) storage.mode(v) <- "double"
      e <- raster::extent(r)
      list(values = v, dim = c(ncol(r), nrow(r), raster::nlayers(r)),
           geo = c(e@xmin, e@ymax, raster::xres(r), raster::yres(r)),
           divisor = if (n %in% names(divisor)) as.double(divisor[[n]]) else NULL)
    } else as.double(r)
  })
  fix.idx <- fix.delta <- NULL
  if (length(cov.fix)) {
    fix.idx <- t(vapply(cov.fix, function(x) match(as.character(x[1:2]), covn), c(1L, 1L)))
    if (anyNA(fix.idx)) stop("pred.maps: cov.fix names a covariate that was not supplied")
    fix.delta <- vapply(cov.fix, function(x) as.double(x[[3]]), 1)
  }
  same <- identical(gg, gg2)
  res <- .Call(C_rfsi_pred_maps, fo$handle, as.matrix(gg2) + 0, if (same) NULL else as.matrix(gg) + 0,
               as.matrix(obs.xy) + 0, as.matrix(obs.z) + 0, as.integer(n.obs), cov, as.integer(code),
               if (is.null(quantiles)) NULL else as.double(quantiles), fix.idx, fix.delta, as.integer(device))
  names(res) <- c("pred", "iqr")
  res
}
