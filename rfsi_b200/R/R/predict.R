# predict for ranger regression forests on the engine -- predict.ranger's call shapes as the
# scripts use them (simulation/simulation.R:214; sim_500.R:156; temp_croatia/2_predictions.R:167,174-176;
# prcp_catalonia/5_prcp_prediction.R:287-291): returns a list with $predictions.
#
# predict.ranger() below MASKS ranger's own S3 method once this package is loaded after ranger
# (NAMESPACE registers it; the later registration wins), so `predict(rfsi_model, df)` and
# `predict(rfsi_model, df, type = "quantiles", quantiles = c(.25, .5, .75))` in the scripts land
# in the CUDA engine with no edit beyond `library(rfsib200)`.  Anything the engine does not
# cover (classification / survival forests, predict.all, se, a subset of trees, factor splits)
# is handed to ranger's own method unchanged.

.rfsi_cache <- new.env(parent = emptyenv())

# the engine handle of a ranger model, imported once per model and R session
rfsi_forest_of <- function(model) {
  if (inherits(model, "rfsi_forest")) return(model)
  key <- paste(model$num.trees, format(model$prediction.error, digits = 17), format(model$r.squared, digits = 17),
               length(model$forest$split.values[[1]]), sum(model$forest$split.values[[1]]), sep = "|")
  fo <- .rfsi_cache[[key]]
  if (is.null(fo)) {
    fo <- rfsi_import_ranger(model)
    assign(key, fo, envir = .rfsi_cache)
  }
  fo
}

predict.rfsi_forest <- function(object, data, type = c("response", "quantiles", "terminalNodes"),
                                quantiles = c(0.1, 0.5, 0.9), device = 0L, ...) {
  type <- match.arg(type)
  vars <- object$independent.variable.names
  miss <- setdiff(vars, names(data))
  if (length(miss)) stop("Error: One or more independent variables not found in data: ", paste(miss, collapse = ", "))
  X <- data.matrix(as.data.frame(data)[, vars, drop = FALSE])   # by name; extra columns ignored (simulation.R:214)
  storage.mode(X) <- "double"
  if (type == "response") {
    p <- .Call(C_rfsi_predict, object$handle, X, NULL, as.integer(device))
  } else if (type == "quantiles") {
    if (!object$quantreg) stop("Error: Set quantreg=TRUE in ranger(...) for quantile prediction.")
    p <- .Call(C_rfsi_predict, object$handle, X, as.double(quantiles), as.integer(device))
    colnames(p) <- paste("quantile=", quantiles)
  } else {
    p <- .Call(C_rfsi_terminal_nodes, object$handle, X, as.integer(device))
  }
  structure(list(predictions = p, num.trees = object$num.trees, num.independent.variables = length(vars),
                 num.samples = nrow(X), treetype = "Regression"), class = "ranger.prediction")
}

predict.ranger <- function(object, data = NULL, predict.all = FALSE, num.trees = object$num.trees, type = "response",
                           se.method = "infjack", quantiles = c(0.1, 0.5, 0.9), what = NULL, seed = NULL,
                           num.threads = NULL, verbose = TRUE, ...) {
  fo <- object$forest
  on.engine <- !is.null(fo) && identical(object$treetype, "Regression") && !predict.all && is.null(what) &&
    num.trees == object$num.trees && type %in% c("response", "quantiles", "terminalNodes") &&
    (is.null(fo$is.ordered) || all(fo$is.ordered)) && !is.null(data)
  if (!on.engine) {
    own <- get("predict.ranger", envir = asNamespace("ranger"))
    return(own(object, data = data, predict.all = predict.all, num.trees = num.trees, type = type,
               se.method = se.method, quantiles = quantiles, what = what, seed = seed,
               num.threads = num.threads, verbose = verbose, ...))
  }
  predict.rfsi_forest(rfsi_forest_of(object), data, type = type, quantiles = quantiles, ...)
}

# explicit form, for callers who do not want the mask
predict_b200 <- function(object, data, ...) predict.rfsi_forest(rfsi_forest_of(object), data, ...)
