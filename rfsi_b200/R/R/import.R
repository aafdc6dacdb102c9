# rfsi_import_ranger -- flatten a ranger regression forest (the object saved at
# temp_croatia/1_models.R:1056-1063 and loaded at temp_croatia/2_predictions.R:59) into the
# contiguous vectors rfsi_forest_create() takes, and return an engine handle.
rfsi_import_ranger <- function(model) {
  stopifnot(inherits(model, "ranger"), model$treetype == "Regression")
  fo <- model$forest
  if (!is.null(fo$is.ordered) && !all(fo$is.ordered)) stop("unordered (factor) splits are not supported")
  T <- fo$num.trees
  n <- vapply(fo$split.varIDs, length, 1L)
  left <- unlist(lapply(fo$child.nodeIDs, `[[`, 1), use.names = FALSE)
  right <- unlist(lapply(fo$child.nodeIDs, `[[`, 2), use.names = FALSE)
  var <- unlist(fo$split.varIDs, use.names = FALSE)
  val <- unlist(fo$split.values, use.names = FALSE)
  # ranger < 0.11.5 counts the response column in split.varIDs (forest$dependent.varID)
  if (!is.null(fo$dependent.varID)) var <- ifelse(var > fo$dependent.varID, var - 1L, var)
  h <- .Call(C_rfsi_forest_new, as.double(c(0, cumsum(n))), as.integer(left), as.integer(right),
             as.integer(var), as.double(val), model$random.node.values,
             length(fo$independent.variable.names))
  structure(list(handle = h, num.trees = T, independent.variable.names = fo$independent.variable.names,
                 quantreg = !is.null(model$random.node.values)), class = "rfsi_forest")
}
